#!/bin/bash
# round 2, GPU contact F (8 GPUs): C5 sweep 4096 configs x 1e7 packets config-per-GPU,
# mccyl variant, C3 end-to-end scaling, 2-GPU NCCL tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --config c5_slab --sweep 512 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_sweep4096_n8.json 2> gpurun_out/r2f_sweep4096_n8.err
tail -c 700 gpurun_out/r2f_sweep4096_n8.json; tail -2 gpurun_out/r2f_sweep4096_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --config c5_cyl --sweep 64 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_sweepcyl512_n8.json 2> gpurun_out/r2f_sweepcyl512_n8.err
tail -c 500 gpurun_out/r2f_sweepcyl512_n8.json; tail -2 gpurun_out/r2f_sweepcyl512_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --config c3_vox --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_c3_n8.json 2> gpurun_out/r2f_c3_n8.err
tail -c 900 gpurun_out/r2f_c3_n8.json; tail -2 gpurun_out/r2f_c3_n8.err
timeout 300 python bench.py --config c5_slab --sweep 512 --packets 1e7 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_sweep512_n1.json 2> gpurun_out/r2f_sweep512_n1.err
tail -c 500 gpurun_out/r2f_sweep512_n1.json
timeout 300 python bench.py --config c3_vox --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_c3_n1.json 2> gpurun_out/r2f_c3_n1.err
tail -c 700 gpurun_out/r2f_c3_n1.json
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3
