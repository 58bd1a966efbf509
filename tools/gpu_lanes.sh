#!/bin/bash
mkdir -p gpurun_out
T=r03p
timeout 500 python -m pytest tests/test_sweep.py tests/test_multi_gpu.py tests/test_gpu_validate.py -m gpu -q -x --timeout=200 2>&1 | tail -6 | tee gpurun_out/${T}_pytest_sweep.log
timeout 600 python bench.py --config c5_slab --sweep 64 --packets 1e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_c5_sweep64.json 2> gpurun_out/${T}_bench_c5_sweep64.err
timeout 600 python bench.py --config validate_uniformfiber --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_validate.json 2> gpurun_out/${T}_bench_validate.err
timeout 600 python bench.py --config c5_cyl --sweep 16 --packets 1e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_c5cyl_sweep16.json 2> gpurun_out/${T}_bench_c5cyl_sweep16.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03p_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_')[1], 'value %.4e e2e %.4e ms/step %.2f frac %.3f vs_baseline %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d.get('vs_baseline')))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1200:])
P
