mkdir -p gpurun_out
T=r05f
for c in c4_trace c4_trace_vox; do
timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; tail -2 gpurun_out/${T}_bench_$c.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r05f_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.4e e2e %.4e chk %.6e'%(d['value'], d['e2e']['value'], d['e2e']['checksum']), d['e2e'])
P
