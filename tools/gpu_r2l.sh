#!/bin/bash
# round 2, GPU contact L: binary64 kernels against the reference kernel's golden vectors
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_double_precision.py tests/test_sweep.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-260 | head -40
