#!/bin/bash
# user surface layouts on the GPU, both mcvox loops, ncu of the voxel trace configuration
mkdir -p gpurun_out
T=r02x
timeout 900 python -m pytest tests -m gpu -q -x -k "user or both_mcvox" > gpurun_out/${T}_pytest_user.log 2>&1; tail -6 gpurun_out/${T}_pytest_user.log
timeout 300 python tools/probe_config.py c4_trace_vox 2e5 2>&1 | tee gpurun_out/${T}_probe_c4vox.log
timeout 600 tools/gpu_ncu.sh c4_trace_vox 2e5 $T
timeout 300 python tools/probe_config.py c4_trace 1e6 2>&1 | tee gpurun_out/${T}_probe_c4.log
