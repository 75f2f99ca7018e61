#!/bin/bash
# per-kernel launch list of three eager SVGD epochs (scripts/svgd_one.py) under ncu, after the same command ran clean without it
mkdir -p gpurun_out
TAG=${TAG:-r02}
for n in ${NS:-4096}; do
CMD="python scripts/svgd_one.py $n ${WL:-W4}"
timeout 200 $CMD > gpurun_out/${TAG}_svgd_plain_$n.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_svgd_launches_$n.csv $CMD > gpurun_out/${TAG}_svgd_ncu_launches_$n.log 2>&1
tail -1 gpurun_out/${TAG}_svgd_plain_$n.log
done
