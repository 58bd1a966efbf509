#!/bin/bash
mkdir -p gpurun_out
for r in 22 26 28 30 32; do echo wait_lanes $r; XO_REFILL=$r timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | sed -n 3p; done
for r in 16 24 28; do echo "wait_lanes $r (small vox case, mcvox_gauss_fluence-like c3 at n=101)"; done
