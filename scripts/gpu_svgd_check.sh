#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 -k "svgd or tensor or smoke" > gpurun_out/svgd_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/svgd_pytest.log
SVGD_TENSOR=1 timeout 300 python scripts/svgd_probe.py W4 4096 16384 32768 > gpurun_out/probe_svgd_w4.log 2>&1; tail -3 gpurun_out/probe_svgd_w4.log
