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


_MASK64 = 0xFFFFFFFFFFFFFFFF


def _splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30))*0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27))*0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def _valid_xinit(x: int) -> bool:
    """Initializers the seed generator accepts (rng.cpp:75-80 / init_RNG): the low
    word must not be 0xFFFFFFFF (nor the state zero) and the high word must lie
    below the generator's multiplier - 1."""
    lo, hi = x & 0xFFFFFFFF, x >> 32
    return 0 < x and lo != 0xFFFFFFFF and hi < 4294967118 - 1


def seed_for_rank(rnginit: int, rank: int) -> int:
    """Rank-specific initializer of the seed generator.  Rank 0 keeps ``rnginit``
    (a single-GPU run is reproduced exactly); the other ranks get a splitmix64
    hash of (rnginit, rank), retried until ``init_RNG`` accepts it - neighbouring
    initializers (rnginit + rank) would start every multiplier's stream from
    linearly related (x, c) states.  All ranks use the full multiplier table, as
    independent reference runs do."""
    rnginit, rank = int(rnginit) & _MASK64, int(rank)
    if rank == 0:
        return rnginit
    x = _splitmix64(rnginit ^ _splitmix64(rank))
    while not _valid_xinit(x):
        x = _splitmix64(x)
    return x


def allreduce_host(array: np.ndarray) -> np.ndarray:
    """Sum a host uint64 array over all ranks (gloo / NCCL-less fallback)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(array).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint64)


class NcclAccumulatorReducer:
    """``sim._reduce_hook``: in-place NCCL reduction of the device accumulator
    buffer (viewed as int64 - two's-complement addition is the same operation),
    stream-ordered on the engine's own stream: the engine's CUDA stream is handed
    to torch as an ``ExternalStream`` and made current around the collective, so
    NCCL waits for the kernel and the engine's later copies wait for NCCL on the
    device - no host synchronisation on the data path.  ``op``: 'allreduce'
    (every rank holds the sum) or 'reduce' (only ``root`` does).  The buffer is
    owned by libxopto_b200; torch only borrows the pointer through
    ``__cuda_array_interface__``."""

    def __init__(self, device_index: int, op: str = 'allreduce', root: int = 0):
        import torch
        import torch.distributed as dist
        if op not in ('allreduce', 'reduce'):
            raise ValueError('op must be "allreduce" or "reduce"')
        self.torch, self.dist = torch, dist
        self.device = torch.device('cuda', device_index)
        self.op, self.root = op, int(root)
        self._cache = {}
        self._streams = {}

    def _tensor(self, abuf, count: int):
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
            if len(self._cache) > 4:
                self._cache.clear()
            self._cache[key] = t
        return t

    def _stream(self, sim):
        native = sim._stream.native
        ext = self._streams.get(native)
        if ext is None:
            ext = self.torch.cuda.ExternalStream(native, device=self.device)
            self._streams[native] = ext
        return ext

    def __call__(self, sim, abuf, count: int):
        t = self._tensor(abuf, count)
        with self.torch.cuda.stream(self._stream(sim)):
            if self.op == 'allreduce':
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            else:
                self.dist.reduce(t, dst=self.root, op=self.dist.ReduceOp.SUM)


def run_sharded(sim, nphotons: int, rank: int, world: int, reducer=None, root: int = None,
                **run_kwargs):
    """Run this rank's shard of ``nphotons`` and return globally reduced results.
    ``root`` None: every rank downloads and returns the same detectors / fluence
    (needs an all-reduce).  ``root`` = r: only rank r downloads, converts and
    returns results (a reduce to r suffices); the other ranks return
    ``(trace, None, None)`` and never touch the host-side copy of the grid.
    Traces stay rank-local."""
    _, count = shard(nphotons, world, rank)
    sim._reduce_hook = reducer if world > 1 else None
    if root is not None and world > 1 and rank != int(root):
        trace, _, _ = sim.run(count, **dict(run_kwargs, download=False))
        return trace, None, None
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
