#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "grid_conversion or fast_mode or profiles" 2>&1 | tail -4
timeout 600 python bench.py --config c3_vox --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], 'e2e', d['e2e'], d['roofline'].get('atomic'))"
timeout 900 python bench.py --config c5_slab --sweep 64 --packets 1e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01u_bench_c5_sweep64_1e7.json 2>gpurun_out/sweep.err; tail -c 1500 gpurun_out/r01u_bench_c5_sweep64_1e7.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sweep', d['value'], 'e2e', d['e2e']['value'], d['roofline']['frac'])"
