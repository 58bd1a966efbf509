#!/bin/bash
mkdir -p gpurun_out
T=r03g
timeout 500 python -m pytest tests -m gpu -q -x --timeout=120 -k "user or sampling_volume or double" 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_sel.log
timeout 300 python bench.py --config c4_trace --steps 5 --warmup 3 > gpurun_out/${T}_bench_c4_trace.json 2> gpurun_out/${T}_bench_c4_trace.err
timeout 300 python bench.py --config c4_trace_vox --steps 5 --warmup 3 > gpurun_out/${T}_bench_c4_trace_vox.json 2> gpurun_out/${T}_bench_c4_trace_vox.err
python - <<'P'
import json
for c in ('c4_trace', 'c4_trace_vox'):
    d=json.loads(open('gpurun_out/r03g_bench_%s.json' % c).read().strip().splitlines()[-1])
    print(c, 'value %.4e e2e %.4e ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
P
