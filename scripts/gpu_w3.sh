#!/bin/bash
for wl in W3; do for cfg in "2 2" "2 1" "3 1" "4 2"; do set -- $cfg
OCBNN_HMC_MB=$1 OCBNN_HMC_U=$2 timeout 200 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']; print('$wl MB=$1 U=$2 value %.4g ms/step %.3f mufu_frac %.3f fp32_frac %.3f accept %.3f mufu/step %d flop/step %d' % (d['value'], d['ms_per_step'], r['mufu_frac'], r['fp32_frac'], d['config']['accept_rate'], r['mufu_per_chain_step'], r['flop_per_chain_step']))"
done; done
