#!/bin/bash
# round 2, GPU contact C: Trace stream as 256-bit stores (C4), parity suite, C2 bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for mb in 3 4; do echo min_blocks $mb; XO_MIN_BLOCKS=$mb timeout 300 python tools/probe_config.py c4_trace 1e6 2>&1 | sed -n 2,3p; done
XO_MIN_BLOCKS=4 timeout 300 python tools/probe_config.py c4_trace 1e6 128 2>&1 | sed -n 3p
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --config c4_trace --steps 3 --warmup 3 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
tail -c 1500 gpurun_out/r2c_bench_c4.json; tail -3 gpurun_out/r2c_bench_c4.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
tail -c 900 gpurun_out/r2c_bench_c2.json; tail -3 gpurun_out/r2c_bench_c2.err
timeout 600 tools/gpu_ncu.sh c4_trace 1e6 r02c
