#!/bin/bash
# round 2, GPU contact B: C2 loop after the XU-pipe relief (no F2I in the deposit, one
# MUFU less in the rotation, MHg with two draws): timing, refill sweep, statistics, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 tools/trace_store_probe.bin > gpurun_out/trace_store_probe_r02.json 2> gpurun_out/trace_store_probe_r02.err; grep -E "fill_fma\": (100|200)" gpurun_out/trace_store_probe_r02.json | cut -c1-120
for c in "c2_skin 1.25e8" "c1_slab 1e8" "c5_slab 1e7"; do
  timeout 300 python tools/probe_config.py $c 2>&1 | tail -3
done
for r in 4 8 12 16; do echo refill $r; XO_REFILL=$r timeout 300 python tools/probe_config.py c2_skin 1.25e8 2>&1 | sed -n 3p; done
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py tests/test_gpu_validate.py -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
timeout 600 tools/gpu_ncu.sh c2_skin 2e7 r02b
