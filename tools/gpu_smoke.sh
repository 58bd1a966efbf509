#!/bin/bash
# GPU contact: device info, parity tests, quick timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
head -c 6000 gpurun_out/pytest_gpu.log
echo ...
tail -n 25 gpurun_out/pytest_gpu.log
