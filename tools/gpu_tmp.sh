mkdir -p gpurun_out
T=r05d
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -X faulthandler -m pytest tests -m gpu -q --timeout=240 --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -9 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -2 gpurun_out/${T}_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_default.json 2> gpurun_out/${T}_bench_reference_default.err
python - <<'P'
import json
for f in ('gpurun_out/r05d_bench_default.json', 'gpurun_out/r05d_bench_reference_default.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); s=d.get('secondary')
    print(f, 'value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), ('| C3 %.4e e2e %.4e frac %.3f'%(s['value'], s['e2e']['value'], s['roofline']['frac'])) if s else '')
P
