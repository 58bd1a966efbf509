#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | sed -n 3p
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_headline.py -m gpu -q -k "reference_kernel and c3_vox" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-300 | head -8; done
timeout 600 python -m pytest tests/test_sweep.py -m gpu -q 2>&1 | tail -3
