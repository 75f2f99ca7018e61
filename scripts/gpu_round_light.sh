#!/bin/bash
# gpu_round.sh without the ncu passes: GPU parity tests, smoke, default bench line, reference arm, extras, W6 / W7 probes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 400 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extras > gpurun_out/bench_extras.log 2> gpurun_out/bench_extras.err; tail -c 400 gpurun_out/bench_extras.log
SVGD_TENSOR=1 timeout 200 python scripts/svgd_probe.py W7 1024 4096 > gpurun_out/probe_svgd_w7.log 2>&1; tail -2 gpurun_out/probe_svgd_w7.log
timeout 200 python scripts/bbb_probe.py 4096 > gpurun_out/probe_bbb_w6.log 2>&1; tail -1 gpurun_out/probe_bbb_w6.log
