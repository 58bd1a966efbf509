#!/bin/bash
mkdir -p gpurun_out
for h in 0 1 2; do
  XOPTO_TRACE_STORE_HINT=$h timeout 200 python tools/probe_config.py c4_trace_vox 2e5 2>&1 | grep kernel | tail -1 | sed "s/^/hint $h: /" | tee -a gpurun_out/r03l_store_hint.log
  XOPTO_TRACE_STORE_HINT=$h timeout 200 python tools/probe_config.py c4_trace 1e6 2>&1 | grep kernel | tail -1 | sed "s/^/hint $h: /" | tee -a gpurun_out/r03l_store_hint.log
done
