mkdir -p gpurun_out
T=r05h
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_default_n8.json 2> gpurun_out/${T}_bench_default_n8.err; tail -3 gpurun_out/${T}_bench_default_n8.err
python - <<'P'
import json
for f in ('gpurun_out/r05h_bench_default_n8.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); s=d.get('secondary')
        print(f, d.get('n_gpus'), 'value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), ('| C3 %.4e e2e %.4e'%(s['value'], s['e2e']['value'])) if s else '')
    except Exception as e:
        print(f, 'ERR', e)
P
