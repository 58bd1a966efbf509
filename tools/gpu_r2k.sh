#!/bin/bash
# round 2, GPU contact K: full parity suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1800 python -X faulthandler -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2k_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2k_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2k_bench_default.json').read().strip().splitlines()[-1])
    s=d['secondary']
    print('C2 value %.4e e2e %.4e frac %.3f | C3 value %.4e e2e %.4e frac %.3f'%(d['value'],d['e2e']['value'],d['roofline']['frac'],s['value'],s['e2e']['value'],s['roofline']['frac']))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2k_bench_default.err').read()[-2000:])
P
