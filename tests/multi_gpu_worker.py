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
    # config-per-GPU sweep: pilot costs exchanged over NCCL, longest-first
    # deal, one all-gather of the rows; every rank ends up with the rows of one GPU
    # simulating all configurations alone (deterministic mode, same seeds per simulator)
    from pyxopto_b200 import mcsweep
    g = 0.8
    configs = [{1: {'mua': float(mua), 'mus': float(musr/(1.0 - g))}}
               for mua in np.linspace(0.0, 5e2, 3) for musr in np.linspace(5e2, 35e2, 3)]

    def sweep_sim():
        return benchcfg.c5_slab(mc, options=opts, cl_devices=local)

    costs = mcsweep.Sweep(sweep_sim(), rank, world).pilot_costs(configs, 500)
    sweep = mcsweep.Sweep(sweep_sim(), rank, world)     # (the pilot advanced its simulator's MWC states)
    idx, rows = sweep.run(configs, 20000, maxthreads=1024, wgsize=64, costs=costs)
    full = sweep.gather(idx, rows, len(configs))
    sweep_ok = full.shape[0] == len(configs) and bool((full.sum(axis=1) > 0).all()) and \
        idx.tolist() == mcsweep.balanced_partition(costs, world, rank).tolist()
    if rank == 0:
        # each configuration starts from the simulator's initial MWC state only when it
        # is the first one of its rank: compare the first configuration of every rank
        for r in range(world):
            first = int(mcsweep.balanced_partition(costs, world, r)[0])
            alone = mcsweep.Sweep(sweep_sim())
            alone._forced = np.array([first])
            _, row = alone.run(configs, 20000, maxthreads=1024, wgsize=64)
            sweep_ok = sweep_ok and bool(np.array_equal(row[0], full[first]))
        print('MULTI_GPU_SWEEP', 'OK' if sweep_ok else 'MISMATCH', flush=True)
    ok = ok and sweep_ok
    flag = torch.tensor([int(ok)], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
    main()
