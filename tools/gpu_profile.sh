#!/bin/bash
# smoke + bench + ncu launch list + one full ncu capture of the top kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --packets 2e7 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/launches_c2.csv
ncu --set full --clock-control none --import-source on -k regex:McKernel -s 1 -c 1 -f -o gpurun_out/prof_c2 \
    python tools/qb.py c2_skin 2e7 > gpurun_out/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:McKernel -s 1 -c 1 -f -o gpurun_out/prof_c1 \
    python tools/qb.py c1_slab 5e6 > gpurun_out/ncu_full_c1.log 2>&1
ls -la gpurun_out
