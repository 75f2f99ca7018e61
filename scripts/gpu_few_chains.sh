#!/bin/bash
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hmc or k2" 2>&1 | tail -4
for lanes in 0 4 8 16; do
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd --chains 4096 --lanes $lanes 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('chains 4096 lanes $lanes  value %.4g  ms/step %.4f mufu_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['config']['accept_rate']))"
done
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd --chains 32768 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('chains 32768 auto  value %.4g  ms/step %.4f mufu_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac']))"
