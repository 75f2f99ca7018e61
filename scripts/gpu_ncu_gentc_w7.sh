#!/bin/bash
mkdir -p gpurun_out
CMD="env SVGD_TENSOR=1 python scripts/svgd_one.py 1024 W7"
timeout 200 $CMD > gpurun_out/gentc_w7_plain.log 2>&1 && tail -1 gpurun_out/gentc_w7_plain.log &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_logpost_gen -s 1 -c 1 -f -o gpurun_out/prof_r02_gentc_w7 $CMD > gpurun_out/ncu_full_gentc_w7.log 2>&1; tail -2 gpurun_out/ncu_full_gentc_w7.log
