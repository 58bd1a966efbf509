#!/bin/bash
mkdir -p gpurun_out
for h in 256 4 6 10; do echo hop_min $h; XOPTO_VOX_HOP_MIN=$h timeout 300 python tools/probe_config.py c3_vox 1e8 2>&1 | sed -n 3p; done
XOPTO_VOX_HOP_MIN=256 timeout 600 python -m pytest tests -m gpu -q -k "mcvox or c3_vox or vox" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q -k "mcvox or c3_vox or vox" 2>&1 | tail -3
