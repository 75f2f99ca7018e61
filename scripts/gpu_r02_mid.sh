#!/bin/bash
# full GPU parity suite + smoke + the application-shape bench lines with the tcgen05 K1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --workload W6 > gpurun_out/bench_W6.log 2> gpurun_out/bench_W6.err; tail -c 400 gpurun_out/bench_W6.log
timeout 400 python bench.py --workload W7 > gpurun_out/bench_W7.log 2> gpurun_out/bench_W7.err; tail -c 400 gpurun_out/bench_W7.log
