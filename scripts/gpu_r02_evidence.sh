#!/bin/bash
# Round-2 evidence: GPU parity tests, smoke, default bench line, then ncu: launch lists of the bench command and of three eager
# SVGD epochs, one full capture each of k_hmc_small, k_phi_flash and k_count_tc.  Every step under its own timeout; ncu only after
# the same command exited 0 without it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 500 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hmc_small -s 4 -c 1 -f -o gpurun_out/prof_r02_hmc $CMD > gpurun_out/ncu_full_hmc.log 2>&1; tail -1 gpurun_out/ncu_full_hmc.log
CMD="python scripts/svgd_one.py 4096 W4"
timeout 200 $CMD > gpurun_out/r02_svgd_plain_4096.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_svgd_launches_4096.csv $CMD > gpurun_out/r02_svgd_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_phi_flash -s 1 -c 1 -f -o gpurun_out/prof_r02_phi_flash $CMD > gpurun_out/ncu_full_phi.log 2>&1; tail -1 gpurun_out/ncu_full_phi.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_count_tc -s 1 -c 1 -f -o gpurun_out/prof_r02_count_tc $CMD > gpurun_out/ncu_full_count.log 2>&1; tail -1 gpurun_out/ncu_full_count.log
CMD="python scripts/svgd_one.py 32768 W4"
timeout 200 $CMD > gpurun_out/r02_svgd_plain_32768.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:k_phi_flash -s 1 -c 1 -f -o gpurun_out/prof_r02_phi_flash_32768 $CMD > gpurun_out/ncu_full_phi32k.log 2>&1; tail -1 gpurun_out/ncu_full_phi32k.log
SVGD_TENSOR=1 timeout 300 python scripts/svgd_probe.py W4 4096 16384 32768 > gpurun_out/probe_svgd_w4.log 2>&1; tail -3 gpurun_out/probe_svgd_w4.log
ls -la gpurun_out | head -40
