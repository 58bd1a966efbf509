#!/bin/bash
# experiment: service-round mcml loop -- parity subset + threshold sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput or scale or deterministic_mode" 2>&1 | tail -6
timeout 300 python tools/exp_refill.py c2_skin 4e7 2,3,4,5,6,8,10 2>&1 | tail -9
timeout 300 python tools/exp_refill.py c1_slab 1e7 1,2,3,4,6,8 2>&1 | tail -8
timeout 200 python tools/exp_refill.py c5_slab 1e7 1,2,4,6 2>&1 | tail -6
