#!/bin/bash
# weak-scaling lines on N GPUs of one box: tools/gpu_scale.sh N TAG   (run under gpurun --gpus N)
N=${1:-2}; TAG=${2:-r01z}
mkdir -p gpurun_out
for c in c2_skin c3_vox; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_scale_n${N}_$c.json 2> gpurun_out/scale_n${N}.err
  tail -c 900 gpurun_out/${TAG}_scale_n${N}_$c.json; tail -2 gpurun_out/scale_n${N}.err
done
