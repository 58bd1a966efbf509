#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput or scale" --durations=8 2>&1 | tail -16
timeout 300 python tools/exp_knobs.py c2_skin 4e7 "pop_batch=1" 2>&1 | tail -1
