#!/bin/bash
# Round evidence: GPU parity tests, smoke, the default bench line (with cpu_baseline + e2e), the reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.log
TAG=hmc_v2 bash scripts/gpu_ncu.sh > gpurun_out/ncu_script.log 2>&1; tail -3 gpurun_out/ncu_script.log
