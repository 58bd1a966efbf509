#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout=300 2>&1 | tail -40 | tee gpurun_out/r03h_pytest_multi_gpu.log
