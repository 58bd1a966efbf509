#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/c4_e2e_profile.py c4_trace 1e6 > gpurun_out/r03d_c4_e2e_profile.txt 2>&1; head -80 gpurun_out/r03d_c4_e2e_profile.txt | cut -c1-150
