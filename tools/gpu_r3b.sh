#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout=90 -k "sampling_volume or sv" 2>&1 | tail -8 | tee gpurun_out/r03b_pytest_sv.log
timeout 200 python tools/c4_stages.py c4_trace 1e6 2>&1 | tail -2 | tee gpurun_out/r03b_c4_stages.log
timeout 200 python tools/c4_stages.py c4_trace_vox 2e5 2>&1 | tail -2 | tee -a gpurun_out/r03b_c4_stages.log
