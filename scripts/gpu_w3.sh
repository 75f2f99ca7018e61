#!/bin/bash
# HMC on the classifier workload W3 (F2a: [2,10,3] + Dirichlet COCP) and the vanilla W1: kernel variants
for cfg in "default" "2 2"; do
if [ "$cfg" = "default" ]; then ENV="X=1"; else set -- $cfg; ENV="OCBNN_HMC_MB=$1 OCBNN_HMC_U=$2"; fi
env $ENV timeout 200 python bench.py --workload W3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']; print('W3 [$cfg] value %.4g ms/step %.3f mufu_frac %.3f fp32_frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], r['mufu_frac'], r['fp32_frac'], d['config']['accept_rate']))"
done
timeout 200 python bench.py --workload W1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']; print('W1 value %.4g ms/step %.3f mufu_frac %.3f fp32_frac %.3f' % (d['value'], d['ms_per_step'], r['mufu_frac'], r['fp32_frac']))"
