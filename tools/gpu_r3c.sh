#!/bin/bash
mkdir -p gpurun_out
T=r03c
for c in c4_trace c4_trace_vox; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r03c_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f.split('bench_')[1], 'value %.4e e2e %.4e ms/step %.3f kernel %.3f frac %.3f share %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['kernel_share_of_step']))
P
