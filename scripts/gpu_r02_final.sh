#!/bin/bash
# Round-2 closing evidence on one B200: GPU parity suite, smoke, the default bench line and the reference arm, the W4 / W6 / W7 lines,
# SVGD stage probes, and the ncu passes whose summaries go to profiles/ (each only after the same command ran clean without ncu).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.log
for w in W4 W6 W7; do timeout 400 python bench.py --workload $w > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; tail -c 200 gpurun_out/bench_$w.log; done
SVGD_TENSOR=1 timeout 300 python scripts/svgd_probe.py W4 4096 16384 32768 > gpurun_out/probe_svgd_w4.log 2>&1; tail -3 gpurun_out/probe_svgd_w4.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMD="python scripts/bbb_probe.py 4096"
timeout 200 $CMD > gpurun_out/gentc_plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_logpost_gen -s 1 -c 1 -f -o gpurun_out/prof_r02_gentc_4096 $CMD > gpurun_out/ncu_full_gentc.log 2>&1; tail -1 gpurun_out/ncu_full_gentc.log
CMD="python scripts/svgd_one.py 4096 W4"
timeout 200 $CMD > gpurun_out/r02_svgd_plain_4096.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_svgd_launches_4096.csv $CMD > gpurun_out/r02_svgd_ncu_launches.log 2>&1
CMD="python scripts/svgd_one.py 32768 W4"
timeout 200 $CMD > gpurun_out/r02_svgd_plain_32768.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:k_phi_flash -s 1 -c 1 -f -o gpurun_out/prof_r02_phi_flash_32768 $CMD > gpurun_out/ncu_full_phi32k.log 2>&1; tail -1 gpurun_out/ncu_full_phi32k.log
ls gpurun_out | head -60
