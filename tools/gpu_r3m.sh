#!/bin/bash
# 8 GPUs: default bench line (C2 + C3 secondary) with the final kernels
mkdir -p gpurun_out
T=r03m
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_default_n8.json 2> gpurun_out/${T}_bench_default_n8.err; tail -2 gpurun_out/${T}_bench_default_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r03m_bench_default_n8.json').read().strip().splitlines()[-1]); s=d.get('secondary')
print('n_gpus', d['n_gpus'], 'C2 value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), '| C3 %.4e e2e %.4e'%(s['value'], s['e2e']['value']))
P
