#!/bin/bash
# ncu evidence for the bench command: (1) launch list with device times, (2) one full capture of the dominant kernel.
mkdir -p gpurun_out
CMD="env OCBNN_HMC_MB=${MB:-3} OCBNN_HMC_U=${U:-4} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd --chains ${CHAINS:-262144}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_hmc_small} -s 4 -c 1 -f -o gpurun_out/prof_${TAG:-hmc} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
