#!/bin/bash
for wv in 32 41 42; do
OCBNN_HMC_WSM=$wv timeout 200 python bench.py --workload W3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']; print('W3 [wsm $wv] value %.4g ms/step %.3f fp32_frac %.3f' % (d['value'], d['ms_per_step'], r['fp32_frac']))"
done
OCBNN_HMC_MB=3 OCBNN_HMC_U=2 timeout 200 python bench.py --workload W3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']; print('W3 [reg 3,2] value %.4g ms/step %.3f fp32_frac %.3f' % (d['value'], d['ms_per_step'], r['fp32_frac']))"
