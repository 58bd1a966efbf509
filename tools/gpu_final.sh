#!/bin/bash
# round 2, final single-GPU evidence: parity suite, bench lines of every configuration, launch
# list of the default bench command, full ncu captures of the final kernels
mkdir -p gpurun_out
T=${1:-r04b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q --timeout=240 --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -9 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -2 gpurun_out/${T}_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_default.json 2> gpurun_out/${T}_bench_reference_default.err
for c in c1_slab c4_trace c4_trace_vox c5_cyl validate_uniformfiber; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
timeout 600 python bench.py --config c5_slab --sweep 64 --packets 1e7 --steps 2 --warmup 1 > gpurun_out/${T}_bench_c5_sweep64.json 2> gpurun_out/${T}_bench_c5_sweep64.err
T=$T python - <<'P'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % os.environ['T'])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get('secondary')
        print(f.split('bench_')[1], 'value %.4e e2e %.4e frac %.3f'%(d['value'], d['e2e']['value'], (d.get('roofline') or {}).get('frac') or 0), ('| C3 %.4e e2e %.4e frac %.3f'%(s['value'], s['e2e']['value'], (s.get('roofline') or {}).get('frac') or 0)) if s else '')
    except Exception as e:
        print(f, 'ERR', e)
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches_bench_default.log 2>&1
for c in "c2_skin 2e7" "c3_vox 2e7" "c4_trace 1e6"; do
  timeout 600 tools/gpu_ncu.sh $c $T
done
timeout 300 python tools/lazy_fluence_probe.py c3_vox 2e7 5 2>&1 | tail -2 > gpurun_out/${T}_lazy_fluence_probe.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
