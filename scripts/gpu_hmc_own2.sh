#!/bin/bash
mkdir -p gpurun_out
OCBNN_HMC_OWN=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "hmc" > gpurun_out/own2_pytest.log 2>&1; echo "pytest(own=2) rc=$?"; tail -3 gpurun_out/own2_pytest.log
: > gpurun_out/fewchains_own2.log
for chains in 4096 2048; do
for cfg in "1 16" "2 16" "1 32" "2 32"; do
  set -- $cfg
  OCBNN_HMC_OWN=$1 timeout 120 python bench.py --chains $chains --lanes $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>>gpurun_out/fewchains_own2.err | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('chains $chains own-mode $1 lanes $2: value %.4g  ms/step %.4f  mufu_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['config']['accept_rate']))
" >> gpurun_out/fewchains_own2.log
done; done
cat gpurun_out/fewchains_own2.log
