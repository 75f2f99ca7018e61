#!/bin/bash
# ncu evidence for the SVGD tensor path: launch list of one epoch + full capture of the tcgen05 GEMM kernel.
mkdir -p gpurun_out
CMD="env SVGD_TENSOR=1 python scripts/svgd_probe.py W4 4096"
$CMD > gpurun_out/svgd_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/svgd_launches.csv $CMD > gpurun_out/svgd_ncu_launches.log 2>&1
CMD2="env SVGD_TENSOR=1 python scripts/svgd_probe.py W7 2048"
$CMD2 > gpurun_out/svgd_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32x3 -s 2 -c 2 -f -o gpurun_out/prof_svgd_tc $CMD2 > gpurun_out/svgd_ncu_full.log 2>&1
tail -2 gpurun_out/svgd_plain.log; tail -2 gpurun_out/svgd_plain2.log; tail -2 gpurun_out/svgd_ncu_full.log
