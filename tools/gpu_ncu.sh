#!/bin/bash
# one full ncu capture of McKernel for a bench config: tools/gpu_ncu.sh c2_skin 2e7 tag [wgsize]
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:McKernel -s 1 -c 1 -f -o gpurun_out/prof_$3_$1 \
    python tools/probe_config.py $1 $2 $4 > gpurun_out/ncu_full_$3_$1.log 2>&1
tail -3 gpurun_out/ncu_full_$3_$1.log
