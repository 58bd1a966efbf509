"""World-size-2 test of the multi-GPU host logic on CPU (gloo): packet sharding,
rank seeds and the single all-reduce of the integer accumulators.  The compute
of each shard is done by the CPU oracle here (the test-suite may use it; the
product path never does) - what is verified is that shard + all-reduce equals
the sum of the independent shards, exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _shard_accumulators(rank, world, nphotons):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import benchcfg
    import xo_oracle
    from pyxopto_b200 import parallel
    from pyxopto_b200.mcml import mc
    sim = benchcfg.c1_slab(mc, rnginit=parallel.seed_for_rank(123456789, rank))
    first, count = parallel.shard(nphotons, world, rank)
    sim._pack(count)
    desc = xo_oracle.describe(sim, 'mcml')
    res = xo_oracle.run(desc, count, 8, sim.rng_seeds_x[:8], sim.rng_seeds_a[:8])
    return res['accu'], count


def _worker(rank, world, port, nphotons, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from pyxopto_b200 import parallel
    accu, count = _shard_accumulators(rank, world, nphotons)
    total = parallel.allreduce_host(accu)
    np.save(os.path.join(out_dir, 'reduced_{}.npy'.format(rank)), total)
    np.save(os.path.join(out_dir, 'local_{}.npy'.format(rank)), accu)
    dist.destroy_process_group()


def test_shard_partition_is_disjoint_and_covering():
    sys.path.insert(0, ROOT)
    from pyxopto_b200 import parallel
    for n in (0, 1, 7, 1000, 10**9 + 3):
        for world in (1, 2, 3, 8):
            pos = 0
            for rank in range(world):
                first, count = parallel.shard(n, world, rank)
                assert first == pos
                pos += count
            assert pos == n
    seeds = {parallel.seed_for_rank(0x2545F4914F6CDD1D, r) for r in range(8)}
    assert len(seeds) == 8


def test_two_rank_allreduce_equals_sum_of_shards(tmp_path):
    world, nphotons = 2, 3001
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nphotons, str(tmp_path)), nprocs=world, join=True)
    local = [np.load(tmp_path / 'local_{}.npy'.format(r)) for r in range(world)]
    reduced = [np.load(tmp_path / 'reduced_{}.npy'.format(r)) for r in range(world)]
    expect = local[0] + local[1]
    assert np.array_equal(reduced[0], expect)
    assert np.array_equal(reduced[1], expect)
    assert not np.array_equal(local[0], local[1])       # different seed sets
    total = expect.sum()/0x7FFFFF/nphotons
    assert 0.3 < total < 0.5
