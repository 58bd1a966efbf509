#!/bin/bash
# round 2, GPU contact D: deferred 256-bit trace store (C4), full parity suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python tools/probe_config.py c4_trace 1e6 2>&1 | sed -n 2,4p
XO_MIN_BLOCKS=2 timeout 300 python tools/probe_config.py c4_trace 1e6 2>&1 | sed -n 3p
timeout 300 python tools/probe_config.py c4_trace 1e6 128 2>&1 | sed -n 3p
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; tail -7 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --config c4_trace --steps 3 --warmup 3 > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
tail -c 1200 gpurun_out/r2d_bench_c4.json; tail -3 gpurun_out/r2d_bench_c4.err
timeout 600 tools/gpu_ncu.sh c4_trace 1e6 r02d
