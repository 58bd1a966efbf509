"""torchrun worker of tests/test_multi_gpu.py: every rank simulates its shard in
deterministic mode on its own GPU, the uint64 accumulators are combined with one
NCCL all-reduce, and rank 0 re-computes every shard alone to check that the
reduced buffer is exactly the sum of the shards."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import benchcfg
    from pyxopto_b200 import parallel
    from pyxopto_b200.mcbase import mcoptions
    from pyxopto_b200.mcml import mc
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nphotons = 200001
    opts = [mcoptions.McDeterministic.on]

    def shard_sim(r, device):
        return benchcfg.c2_skin(mc, rnginit=parallel.seed_for_rank(benchcfg.RNGINIT, r),
                                options=opts, cl_devices=device)

    sim = shard_sim(rank, local)
    reducer = parallel.NcclAccumulatorReducer(local)
    _, fluence, detectors = parallel.run_sharded(
        sim, nphotons, rank, world, reducer=reducer, maxthreads=4096, wgsize=64)
    reduced, _, _ = sim.download_raw()
    ok = True
    if rank == 0:
        expect = np.zeros_like(reduced)
        for r in range(world):
            s = shard_sim(r, local)
            _, count = parallel.shard(nphotons, world, r)
            s.run(count, maxthreads=4096, wgsize=64, download=False)
            expect += s.download_raw()[0]
        ok = bool(np.array_equal(reduced, expect)) and expect.sum() > 0 and \
            fluence.nphotons == nphotons and detectors.top.nphotons == nphotons
        print('MULTI_GPU_RESULT', 'OK' if ok else 'MISMATCH', 'world', world,
              'accumulator sum', int(reduced.sum()), flush=True)
    flag = torch.tensor([int(ok)], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
    main()
