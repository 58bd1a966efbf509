#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput" 2>&1 | tail -3
for spec in "pop_batch=1" "pop_batch=1,refill_lanes=8"; do
  timeout 300 python tools/exp_knobs.py c2_skin 4e7 "$spec" 2>&1 | tail -1
done
timeout 300 python bench.py --config c3_vox --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'])"
