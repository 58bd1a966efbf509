#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "own_their_pinned or grid_conversion or run_returns or batched" 2>&1 | tail -5
timeout 600 python bench.py --config c3_vox --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_bench_c3.json').read().strip().splitlines()[-1])
print('C3 value %.4e e2e %.4e ms %.2f e2e ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']))
P
