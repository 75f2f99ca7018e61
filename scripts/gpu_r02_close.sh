#!/bin/bash
# last pass of the round: full GPU parity suite, smoke, default bench line (what the driver runs), few-chain default
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.log; echo
timeout 200 python bench.py --chains 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-svgd 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('4096 chains, auto lanes: value %.4g ms/step %.4f mufu_frac %.3f lanes %s' % (d['value'], d['ms_per_step'], d['roofline']['mufu_frac'], d['config'].get('lanes_per_chain')))"
