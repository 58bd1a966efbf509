#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/c4_stages.py c4_trace 1e6 2>&1 | tail -3 | tee gpurun_out/r03a_c4_stages.log
timeout 200 python tools/c4_stages.py c4_trace_vox 2e5 2>&1 | tail -3 | tee -a gpurun_out/r03a_c4_stages.log
