#!/bin/bash
# round 2, GPU contact A: full parity suite (incl. headline / validate tests), atomic
# probe, default bench line (C2 + secondary C3), validate suite, C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -n 40 gpurun_out/r2a_pytest_gpu.log
timeout 120 tools/atomic_probe.bin > gpurun_out/atomic_probe_r02.json 2> gpurun_out/atomic_probe_r02.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
tail -c 1500 gpurun_out/r2a_bench_default.json; tail -3 gpurun_out/r2a_bench_default.err
timeout 600 python bench.py --config validate_uniformfiber --steps 2 --warmup 1 > gpurun_out/r2a_bench_validate.json 2> gpurun_out/r2a_bench_validate.err
tail -c 1200 gpurun_out/r2a_bench_validate.json; tail -3 gpurun_out/r2a_bench_validate.err
timeout 600 python bench.py --config c4_trace --steps 3 --warmup 2 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
tail -c 1200 gpurun_out/r2a_bench_c4.json; tail -3 gpurun_out/r2a_bench_c4.err
