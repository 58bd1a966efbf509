#!/bin/bash
# 2 GPUs: NCCL tests (all-reduce of the shards, sweep gather), SV tests, default bench at N=2, C4 lines
mkdir -p gpurun_out
T=r03f
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 -k "multi_gpu or nccl or sampling_volume" 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_default_n2.json 2> gpurun_out/${T}_bench_default_n2.err; tail -2 gpurun_out/${T}_bench_default_n2.err
timeout 300 python bench.py --config c4_trace --steps 5 --warmup 3 > gpurun_out/${T}_bench_c4_trace.json 2> gpurun_out/${T}_bench_c4_trace.err
python - <<'P'
import json
for f in ['gpurun_out/r03f_bench_default_n2.json', 'gpurun_out/r03f_bench_c4_trace.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); s=d.get('secondary')
        print(f, 'n_gpus', d['n_gpus'], 'value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), ('| C3 %.4e e2e %.4e'%(s['value'], s['e2e']['value'])) if s else '')
    except Exception as e:
        print(f, 'ERR', e)
P
