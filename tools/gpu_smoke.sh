#!/bin/bash
# first GPU contact: device info, parity tests, quick timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
python tools/quick_bench.py mcml_c1_slab 1e7 3 2>&1 | tail -5
python tools/quick_bench.py mcml_mhg_gauss_cart_flurz 1e7 2 2>&1 | tail -5
