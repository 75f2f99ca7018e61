#!/bin/bash
run() { env $1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --chains $2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  $1 chains=$2 value %.3e  ms/step %.3f  mufu_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac']))
    except Exception as e:
        print(l.strip()[:200])
"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "hmc or k1" 2>&1 | tail -2
OCBNN_HMC_STATE=smem OCBNN_HMC_MB=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "hmc" 2>&1 | tail -2
for ch in 262144 4096; do
run "OCBNN_HMC_STATE=reg OCBNN_HMC_MB=2" $ch
run "OCBNN_HMC_STATE=smem OCBNN_HMC_MB=3" $ch
run "OCBNN_HMC_STATE=smem OCBNN_HMC_MB=4" $ch
run "OCBNN_HMC_STATE=smem OCBNN_HMC_MB=5" $ch
done
