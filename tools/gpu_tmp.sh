#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/probe_config.py c3_vox 1e6 2>&1 | tail -2
for v in "XO_POOL_SOA=1" "XO_POOL_SOA=0"; do
  env $v timeout 200 python tools/probe_config.py c3_vox 1e8 2>&1 | grep kernel | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/r03t_probe_soa.log
done
timeout 500 python -m pytest tests -m gpu -q -x --timeout=90 -k "vox or c3" 2>&1 | tail -4 | tee gpurun_out/r03t_pytest.log
for v in "XO_POOL_THR_I=16" "XO_POOL_THR_I=24" "XO_POOL_THR_W=8" "XO_POOL_THR_W=16" "XO_POOL_LAUNCH=8" "XO_POOL_LAUNCH=24"; do
  env $v timeout 200 python tools/probe_config.py c3_vox 1e8 2>&1 | grep kernel | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/r03t_probe_soa.log
done
timeout 600 tools/gpu_ncu.sh c3_vox 2e7 r03t
