#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --chains ${CHAINS:-4096} --lanes ${LANES:-16} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd"
timeout 200 $CMD > gpurun_out/fewchains_plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_hmc -s 3 -c 1 -f -o gpurun_out/prof_r02_hmc_fewchains $CMD > gpurun_out/ncu_full_fewchains.log 2>&1; tail -2 gpurun_out/ncu_full_fewchains.log
