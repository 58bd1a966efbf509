#!/bin/bash
# round 2, GPU contact L: binary64 kernels against the reference kernel's golden vectors,
# user-written fluence fragment, sweep partition
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_double_precision.py tests/test_sweep.py -m gpu -q -x 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "user_fragments or batched" 2>&1 | tail -8
