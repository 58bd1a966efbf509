#!/bin/bash
# full parity suite with the packet-pool loop as the mcvox throughput path, default bench line
# (C2 + C3 secondary), launch list, full ncu of C3
mkdir -p gpurun_out
T=r02w
timeout 1800 python -X faulthandler -m pytest tests -m gpu -q --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -9 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -2 gpurun_out/${T}_bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02w_bench_default.json').read().strip().splitlines()[-1])
s=d.get('secondary')
print('C2 value %.4e e2e %.4e frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['frac']), '| C3 %.4e e2e %.4e frac %.3f'%(s['value'], s['e2e']['value'], s['roofline']['frac']))
P
timeout 600 tools/gpu_ncu.sh c3_vox 2e7 $T
