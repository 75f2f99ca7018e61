#!/bin/bash
# tcgen05 generic-net K1: parity tests of every generic-net case, then the W6 / W7 probes with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "k1 or generic or bbb or hmc or sgld or aocp or host_api" > gpurun_out/gentc_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gentc_pytest.log
timeout 200 python scripts/bbb_probe.py 4096 > gpurun_out/gentc_bbb_w6.log 2>&1; tail -2 gpurun_out/gentc_bbb_w6.log
OCBNN_GEN_TC=0 timeout 200 python scripts/bbb_probe.py 4096 > gpurun_out/gentc_bbb_w6_old.log 2>&1; tail -1 gpurun_out/gentc_bbb_w6_old.log
SVGD_TENSOR=1 timeout 200 python scripts/svgd_probe.py W7 4096 > gpurun_out/gentc_w7.log 2>&1; tail -1 gpurun_out/gentc_w7.log
OCBNN_GEN_TC=0 SVGD_TENSOR=1 timeout 200 python scripts/svgd_probe.py W7 4096 > gpurun_out/gentc_w7_old.log 2>&1; tail -1 gpurun_out/gentc_w7_old.log
