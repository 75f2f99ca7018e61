#!/bin/bash
# variant sweep only (no tests) + full default bench line with e2e
mkdir -p gpurun_out
: > gpurun_out/sweep.log
for cfg in ${SWEEP:-"4 2" "4 3" "4 4" "3 4" "3 6"}; do
  set -- $cfg
  echo "MB=$1 U=$2" >> gpurun_out/sweep.log
  OCBNN_HMC_MB=$1 OCBNN_HMC_U=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>>gpurun_out/sweep.err | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('  value %.4g  ms/step %.3f  mufu_frac %.3f fp32_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['roofline']['fp32_frac'], d['config']['accept_rate']))
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>gpurun_out/bench_e2e.err; tail -c 1500 gpurun_out/bench_e2e.log; tail -3 gpurun_out/bench_e2e.err
