#!/bin/bash
# GPU contact: parity tests, smoke, bench lines for the three configs, quick probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
cat MEASURED_PEAKS.json 2>/dev/null
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in c2_skin c1_slab c3_vox c5_cyl; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -c 2500 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err
done
