#!/bin/bash
# voxel trace configuration: occupancy (registers per thread) x waiting-lane threshold
mkdir -p gpurun_out
T=r02y
for mb in 2 3 4; do for rf in 12 16 22; do
  XO_MIN_BLOCKS=$mb XO_REFILL=$rf timeout 200 python tools/probe_config.py c4_trace_vox 2e5 2>&1 | grep kernel | tail -1 | sed "s/^/minblocks $mb refill $rf: /" | tee -a gpurun_out/${T}_probe_c4vox.log
done; done
for wg in 128 512; do
  XO_MIN_BLOCKS=1 timeout 200 python tools/probe_config.py c4_trace_vox 2e5 $wg 2>&1 | grep kernel | tail -1 | sed "s/^/wgsize $wg minblocks 1: /" | tee -a gpurun_out/${T}_probe_c4vox.log
done
XO_MIN_BLOCKS=6 XO_REFILL=16 timeout 200 python tools/probe_config.py c4_trace_vox 2e5 128 2>&1 | grep kernel | tail -1 | sed "s/^/wgsize 128 minblocks 6 refill 16: /" | tee -a gpurun_out/${T}_probe_c4vox.log
