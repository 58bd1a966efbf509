#!/bin/bash
# C3 convergence-barrier experiment: vox parity tests, C3 probe, full ncu of C3 at 2e7 packets
mkdir -p gpurun_out
T=r02s
timeout 900 python -m pytest tests -m gpu -q -x -k "vox or c3" > gpurun_out/${T}_pytest_vox.log 2>&1; tail -4 gpurun_out/${T}_pytest_vox.log
timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | tee gpurun_out/${T}_probe_c3.log
for r in 18 20 24 26; do XO_REFILL=$r timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | grep kernel | tail -1 | sed "s/^/refill $r: /" | tee -a gpurun_out/${T}_probe_c3.log; done
timeout 600 tools/gpu_ncu.sh c3_vox 2e7 $T
