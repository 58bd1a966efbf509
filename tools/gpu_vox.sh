#!/bin/bash
# voxel kernel iteration: statistical parity + timing vs knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput or vox" 2>&1 | tail -5
for spec in "refill_lanes=16,fluence_block=1024" "refill_lanes=16,fluence_block=1024,fluence_window_bytes=0" "refill_lanes=16,fluence_block=1024,fluence_window_bytes=65536" "refill_lanes=12,fluence_block=1024,fluence_window_bytes=0" "refill_lanes=20,fluence_block=1024,fluence_window_bytes=0" "refill_lanes=16,fluence_block=512,fluence_window_bytes=0"; do
  timeout 300 python tools/exp_knobs.py c3_vox 1e7 "$spec" 2>&1 | tail -1
done
