#!/bin/bash
# round-end evidence: tests, smoke, bench lines (all configs + reference arm), ncu
# launch list of the bench command, one full ncu capture per headline kernel
# usage: tools/gpu_final.sh TAG
TAG=${1:-r01z}
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in c2_skin c1_slab c3_vox c5_cyl c4_trace; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/bench_$c.err
  tail -c 600 gpurun_out/${TAG}_bench_$c.json; tail -2 gpurun_out/bench_$c.err
done
timeout 900 python bench.py --config c5_slab --sweep 64 --packets 1e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5_sweep64.json 2> gpurun_out/bench_c5_sweep.err
tail -c 400 gpurun_out/${TAG}_bench_c5_sweep64.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_c2_skin.json 2>&1; tail -c 500 gpurun_out/${TAG}_bench_ref_c2_skin.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c2_skin.csv \
    python bench.py --steps 2 --warmup 1 --packets 2e7 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/${TAG}_launches_bench_c2_skin.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c3_vox.csv \
    python bench.py --config c3_vox --steps 2 --warmup 1 --packets 1e7 --no-cpu-baseline > gpurun_out/bench_under_ncu3.log 2>&1
tail -3 gpurun_out/${TAG}_launches_bench_c3_vox.csv
bash tools/gpu_ncu.sh c2_skin 2e7 $TAG
bash tools/gpu_ncu.sh c3_vox 5e6 $TAG
bash tools/gpu_ncu.sh c1_slab 5e6 $TAG
ls gpurun_out | head -50
