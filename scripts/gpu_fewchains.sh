#!/bin/bash
# few-chain HMC (4096 chains, W2): lanes per chain x points in flight
mkdir -p gpurun_out
: > gpurun_out/fewchains.log
for cfg in ${SWEEP:-"8 4" "16 4" "16 2" "32 2" "32 4" "8 2"}; do
  set -- $cfg
  OCBNN_HMC_GU=$2 timeout 120 python bench.py --chains ${CHAINS:-4096} --lanes $1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>>gpurun_out/fewchains.err | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('chains ${CHAINS:-4096} lanes $1 GU $2: value %.4g  ms/step %.4f  mufu_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['config']['accept_rate']))
" >> gpurun_out/fewchains.log
done
cat gpurun_out/fewchains.log
