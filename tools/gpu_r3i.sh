#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_profile.py c3_vox 1e8 > gpurun_out/r03i_c3_e2e_profile.txt 2>&1; head -50 gpurun_out/r03i_c3_e2e_profile.txt | cut -c1-150
