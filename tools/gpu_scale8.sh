#!/bin/bash
# round 2, 8-GPU evidence: the default bench line (C2 + C3 secondary) at N=8, the 4096-configuration
# sweep with the permutation deal, the mccyl sub-grid sweep
mkdir -p gpurun_out
T=r02n
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_default_n8.json 2> gpurun_out/${T}_bench_default_n8.err
timeout 900 $TR --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/${T}_bench_default_n4.json 2> gpurun_out/${T}_bench_default_n4.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_default_n1.json 2> gpurun_out/${T}_bench_default_n1.err
timeout 900 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --config c5_slab --sweep 512 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_sweep4096_n8.json 2> gpurun_out/${T}_sweep4096_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29534 bench.py --gpus 8 --config c5_cyl --sweep 64 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_sweepcyl512_n8.json 2> gpurun_out/${T}_sweepcyl512_n8.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02n_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get('secondary')
        print(f.split('r02n_')[1], 'N=%d value %.4e e2e %.4e kernel share %.3f'%(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline'].get('kernel_share_of_step') or 0), ('| C3 %.4e e2e %.4e'%(s['value'], s['e2e']['value'])) if s else '')
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
