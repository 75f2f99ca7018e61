#!/bin/bash
mkdir -p gpurun_out
python scripts/bbb_probe.py 1024 > gpurun_out/w6_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_logpost_generic -s 1 -c 1 -f -o gpurun_out/prof_generic_w6 python scripts/bbb_probe.py 1024 > gpurun_out/w6_ncu.log 2>&1
tail -1 gpurun_out/w6_plain.log; tail -2 gpurun_out/w6_ncu.log
SVGD_TENSOR=1 python scripts/svgd_probe.py W7 4096 2>&1 | tail -1
