#!/bin/bash
# GPU box: parity tests, then the HMC kernel variant sweep (OCBNN_HMC_MB x OCBNN_HMC_U) on the bench workload.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.log
for cfg in ${SWEEP:-"4 2" "3 4"}; do
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
