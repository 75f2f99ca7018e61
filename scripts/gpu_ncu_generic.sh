#!/bin/bash
mkdir -p gpurun_out
CMD="env SVGD_TENSOR=1 python scripts/svgd_probe.py W7 1024"
$CMD > gpurun_out/gen_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_logpost_generic -s 2 -c 1 -f -o gpurun_out/prof_generic $CMD > gpurun_out/gen_ncu_full.log 2>&1
tail -1 gpurun_out/gen_plain.log; tail -2 gpurun_out/gen_ncu_full.log
