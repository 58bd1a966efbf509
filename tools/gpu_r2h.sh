#!/bin/bash
# round 2, GPU contact H: mcvox throughput loop (material change inline, cell deposit, cheaper pops)
mkdir -p gpurun_out
timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | sed -n 2,4p
for r in 8 12 20 24; do echo wait_lanes $r; XO_REFILL=$r timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | sed -n 3p; done
timeout 1500 python -m pytest tests -m gpu -q -k "mcvox or c3_vox or sampling_volume or aniso or vox" > gpurun_out/r2i_pytest.log 2>&1; tail -8 gpurun_out/r2h_pytest.log
timeout 600 tools/gpu_ncu.sh c3_vox 5e6 r02i
