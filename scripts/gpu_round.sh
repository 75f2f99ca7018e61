#!/bin/bash
# Round evidence: GPU parity tests, smoke, the default bench line (cpu_baseline + e2e + svgd), the reference arm, extras, probes,
# ncu launch lists + full captures (HMC kernel, fused SVGD phi kernel).  Every step under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 900 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extras > gpurun_out/bench_extras.log 2> gpurun_out/bench_extras.err
SVGD_TENSOR=1 timeout 300 python scripts/svgd_probe.py W4 4096 16384 32768 > gpurun_out/probe_svgd_w4.log 2>&1; tail -3 gpurun_out/probe_svgd_w4.log
bash scripts/gpu_w3.sh > gpurun_out/probe_hmc_w3_w1.log 2>&1; cat gpurun_out/probe_hmc_w3_w1.log
SVGD_TENSOR=1 timeout 300 python scripts/svgd_probe.py W7 1024 4096 > gpurun_out/probe_svgd_w7.log 2>&1; tail -2 gpurun_out/probe_svgd_w7.log
timeout 300 python scripts/bbb_probe.py 4096 > gpurun_out/probe_bbb_w6.log 2>&1; tail -1 gpurun_out/probe_bbb_w6.log
TAG=hmc_v2 timeout 900 bash scripts/gpu_ncu.sh > gpurun_out/ncu_script.log 2>&1; tail -2 gpurun_out/ncu_script.log
CMD="env SVGD_TENSOR=1 python scripts/svgd_probe.py W4 4096"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/svgd_launches_4096.csv $CMD > gpurun_out/svgd_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_phi_fused -s 4 -c 1 -f -o gpurun_out/prof_svgd_fused $CMD > gpurun_out/svgd_ncu_full.log 2>&1; tail -2 gpurun_out/svgd_ncu_full.log
