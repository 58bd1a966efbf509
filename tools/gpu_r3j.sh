#!/bin/bash
mkdir -p gpurun_out
T=r03j
timeout 120 python tools/probe_config.py c4_trace_vox 2e5 2>&1 | tail -4 | tee gpurun_out/${T}_probe_c4vox.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout=120 -k "vox or trace or sampling" 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_sel.log
for v in "XO_POOL_THR_W=8" "XO_POOL_THR_W=16" "XO_POOL_THR_W=20" "XO_POOL_THR_I=8" "XO_POOL_LAUNCH=8" "XO_MIN_BLOCKS=3" "XO_POOL_SLOTS=0"; do
  env $v timeout 200 python tools/probe_config.py c4_trace_vox 2e5 2>&1 | grep kernel | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/${T}_probe_c4vox.log
done
