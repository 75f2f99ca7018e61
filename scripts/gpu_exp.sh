#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for mb in 2 3 4; do
  for ch in 4096 262144; do
    echo "MB=$mb chains=$ch"
    OCBNN_HMC_MB=$mb timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --chains $ch 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('  value %.3e  ms/step %.3f  frac %.3f mufu_frac %.3f accept %.2f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['mufu_frac'], d['config']['accept_rate']))
    except Exception as e:
        print(l.strip()[:300])
"
  done
done
