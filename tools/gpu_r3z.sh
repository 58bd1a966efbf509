#!/bin/bash
# final check of the round: whole GPU suite + default bench line + smoke()
mkdir -p gpurun_out
T=r03z
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q --timeout=240 --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -9 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -2 gpurun_out/${T}_bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r03z_bench_default.json').read().strip().splitlines()[-1]); s=d['secondary']
print('C2 value %.4e e2e %.4e frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['frac']), '| C3 %.4e e2e %.4e frac %.3f'%(s['value'], s['e2e']['value'], s['roofline']['frac']))
P
