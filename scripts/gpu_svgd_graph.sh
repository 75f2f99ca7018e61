#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_parity.py -m gpu -x -q -k "svgd" 2>&1 | tail -4
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['svgd'])"
SVGD_TENSOR=1 timeout 200 python scripts/svgd_probe.py W4 4096 2>&1 | tail -1
