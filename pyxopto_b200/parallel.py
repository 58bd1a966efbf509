"""Multi-GPU sharding of one simulation (the reference has no
equivalent: one ``Mc`` = one OpenCL device, mcworker.py:193-199).

One process per GPU (``torchrun``).  Packets are independent, so rank ``g`` of
``G`` simulates its own slice of the packet budget with a rank-specific MWC seed
set on private buffers - no collective on the data path.  The only exchange is
ONE all-reduce (sum) over the flat uint64 accumulator buffer at the end of
``run()``: integer adds commute, so the result is exactly the sum of the shards.
``torch.distributed`` is plumbing only: NCCL over NVLink on GPUs, gloo in the
CPU test-suite.
"""
import numpy as np


def shard(nphotons: int, world: int, rank: int):
    """(first packet, packet count) of ``rank``: contiguous, disjoint, covering."""
    nphotons, world, rank = int(nphotons), int(world), int(rank)
    q, r = divmod(nphotons, world)
    count = q + (1 if rank < r else 0)
    first = rank*q + min(rank, r)
    return first, count


def seed_for_rank(rnginit: int, rank: int) -> int:
    """Rank-specific initializer of the seed generator.  All ranks use the full
    multiplier table (as independent reference runs do); distinct ``xinit``
    values give distinct (x, c) start states for every multiplier."""
    return (int(rnginit) + int(rank)) & 0xFFFFFFFFFFFFFFFF


def allreduce_host(array: np.ndarray) -> np.ndarray:
    """Sum a host uint64 array over all ranks (gloo / NCCL-less fallback)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(array).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint64)


class NcclAccumulatorReducer:
    """``sim._reduce_hook``: in-place NCCL all-reduce of the device accumulator
    buffer (viewed as int64 - two's-complement addition is the same operation).
    The buffer is owned by libxopto_b200; torch only borrows the pointer through
    ``__cuda_array_interface__``."""

    def __init__(self, device_index: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = torch.device('cuda', device_index)
        self._cache = {}

    def __call__(self, sim, abuf, count: int):
        key = (abuf.device_ptr, int(count))
        t = self._cache.get(key)
        if t is None:
            class _Borrowed:
                pass
            b = _Borrowed()
            b.__cuda_array_interface__ = {
                'shape': (int(count),), 'typestr': '<i8',
                'data': (abuf.device_ptr, False), 'version': 2}
            t = self.torch.as_tensor(b, device=self.device)
            self._cache = {key: t}
        sim._stream.synchronize()        # kernel finished on the engine's stream
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        self.torch.cuda.synchronize(self.device)


def run_sharded(sim, nphotons: int, rank: int, world: int, reducer=None, **run_kwargs):
    """Run this rank's shard of ``nphotons`` and return globally reduced results
    (every rank gets the same detectors/fluence; traces stay rank-local)."""
    _, count = shard(nphotons, world, rank)
    sim._reduce_hook = reducer if world > 1 else None
    trace, fluence, detectors = sim.run(count, **run_kwargs)
    if world > 1:
        # the accumulators now hold the contribution of all packets
        if fluence is not None:
            fluence._nphotons += int(nphotons) - count
        if detectors is not None:
            for det in detectors:
                if hasattr(det, '_nphotons'):
                    det._nphotons += int(nphotons) - count
    return trace, fluence, detectors
