mkdir -p gpurun_out
T=r05g
timeout 600 python -X faulthandler -m pytest tests/test_multi_gpu.py -m gpu -q --timeout=300 > gpurun_out/${T}_pytest_multi_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_multi_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_default_n2.json 2> gpurun_out/${T}_bench_default_n2.err; tail -3 gpurun_out/${T}_bench_default_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_n2.json 2> gpurun_out/${T}_bench_reference_n2.err; tail -3 gpurun_out/${T}_bench_reference_n2.err
python - <<'P'
import json
for f in ('gpurun_out/r05g_bench_default_n2.json','gpurun_out/r05g_bench_reference_n2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); s=d.get('secondary')
        print(f, d.get('n_gpus'), 'value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), ('| C3 %.4e e2e %.4e'%(s['value'], s['e2e']['value'])) if s else '')
    except Exception as e:
        print(f, 'ERR', e)
P
