#!/bin/bash
# voxel kernel iteration: statistical parity + timing vs knobs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput or vox" 2>&1 | tail -5
for spec in "refill_lanes=16" "refill_lanes=12" "refill_lanes=20" "refill_lanes=24"; do
  timeout 300 python tools/exp_knobs.py c3_vox 1e7 "$spec" 2>&1 | tail -1
done

