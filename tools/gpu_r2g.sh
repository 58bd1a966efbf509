#!/bin/bash
# round 2, GPU contact G (2 GPUs): NCCL path (stream-ordered reducer), default bench line at
# N=2, balanced sweep partition, lazy sampling volume + c4_trace_vox
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -25
timeout 900 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_default_n2.json 2> gpurun_out/r2g_default_n2.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2g_default_n2.json').read().strip().splitlines()[-1])
    print('N=2 C2 value %.4e e2e %.4e | C3 value %.4e e2e %.4e'%(d['value'],d['e2e']['value'],d['secondary']['value'],d['secondary']['e2e']['value']))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2g_default_n2.err').read()[-3000:])
P
timeout 600 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --config c5_cyl --sweep 64 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_sweepcyl_n2.json 2> gpurun_out/r2g_sweepcyl_n2.err
tail -c 400 gpurun_out/r2g_sweepcyl_n2.json; tail -2 gpurun_out/r2g_sweepcyl_n2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sampling_volume_accumulates or aniso" 2>&1 | tail -15
timeout 600 python bench.py --config c4_trace --steps 4 --warmup 3 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err
tail -c 1300 gpurun_out/r2g_bench_c4.json | head -c 700; echo; tail -3 gpurun_out/r2g_bench_c4.err
timeout 600 python bench.py --config c4_trace_vox --steps 3 --warmup 2 > gpurun_out/r2g_bench_c4vox.json 2> gpurun_out/r2g_bench_c4vox.err
tail -c 1300 gpurun_out/r2g_bench_c4vox.json | head -c 700; echo; tail -3 gpurun_out/r2g_bench_c4vox.err
