#!/bin/bash
# slot-ownership HMC (k_hmc_own): HMC parity tests, then 4096 / 8192 / 16384 chains at 8 lanes (replicated kernel) and 16 / 32 lanes (ownership)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "hmc" > gpurun_out/own_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/own_pytest.log
CHAINS=4096 SWEEP="8 4|16 2|16 4|32 2|32 4" 
: > gpurun_out/fewchains_own.log
for chains in ${CHAINS_LIST:-4096 8192 16384}; do
for cfg in "8 4" "16 2" "16 4" "32 2" "32 4"; do
  set -- $cfg
  OCBNN_HMC_GU=$2 timeout 120 python bench.py --chains $chains --lanes $1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>>gpurun_out/fewchains_own.err | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('chains $chains lanes $1 GU $2: value %.4g  ms/step %.4f  mufu_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['config']['accept_rate']))
" >> gpurun_out/fewchains_own.log
done; done
cat gpurun_out/fewchains_own.log
