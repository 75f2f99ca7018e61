#!/bin/bash
mkdir -p gpurun_out
for n in ${NS:-4096}; do
CMD="env SVGD_TENSOR=1 python scripts/svgd_probe.py W4 $n"
timeout 200 $CMD > gpurun_out/svgd_plain_$n.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/svgd_launches_$n.csv $CMD > gpurun_out/svgd_ncu_launches_$n.log 2>&1
tail -1 gpurun_out/svgd_plain_$n.log
done
