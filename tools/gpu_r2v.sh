#!/bin/bash
# packet-pool loop (gather by rank), walk loop with pinned strides: parity, probe, sweep, ncu
mkdir -p gpurun_out
T=r02v
timeout 120 python tools/probe_config.py c3_vox 1e6 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q -x -k "vox or c3" > gpurun_out/${T}_pytest_vox.log 2>&1; tail -6 gpurun_out/${T}_pytest_vox.log
timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | tee gpurun_out/${T}_probe_c3.log
for v in "XO_POOL_THR_I=12" "XO_POOL_THR_I=20" "XO_POOL_THR_W=8" "XO_POOL_THR_W=16" "XO_POOL_LAUNCH=8" "XO_POOL_LAUNCH=12" "XO_POOL_BND=8" "XO_POOL_BND=16" "XO_POOL_THR_I=12 XO_POOL_THR_W=8 XO_POOL_LAUNCH=8"; do
  env $v timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | grep kernel | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/${T}_probe_c3.log
done
timeout 600 tools/gpu_ncu.sh c3_vox 2e7 $T
