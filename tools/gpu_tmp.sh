#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/probe_config.py c3_vox 1e8 2>&1 | grep kernel | tail -2 | tee gpurun_out/r03s_probe.log
timeout 400 python -m pytest tests -m gpu -q -x --timeout=90 -k "vox or c3" 2>&1 | tail -4 | tee gpurun_out/r03s_pytest.log
timeout 600 tools/gpu_ncu.sh c3_vox 2e7 r03s
