"""Optical-property sweeps (BASELINE config 5; counterpart of the reference's
sweep callers ``dataset/render/mcml.py:38-95`` + ``mcbase/mcrun.py:49-92``,
which call ``Mc.run`` once per configuration and pay a full host round trip
each time).

A sweep keeps ONE simulator - one compiled kernel, one set of device buffers -
and streams configurations through it:

  * every configuration only rewrites the packed medium / source structs (a few
    hundred bytes) and re-zeroes the accumulators on the device;
  * the accumulator buffer is double-buffered on the device and lands in
    page-locked host rows with asynchronous copies, so configuration i+1 is
    simulated while the results of configuration i cross PCIe;
  * the only synchronisation per configuration is one event wait.

Multi-GPU: configurations are dealt to the ranks by a fixed pseudo-random
permutation (static, the same on every rank), one process per GPU, no collective
on the data path; ``gather()`` assembles the per-rank rows with one all-gather.
(A plain round-robin deal correlates with the layout of a parameter grid: with
(mua, musr) grids of 64 x 64 and 8 ranks every rank got one residue class of the
musr index, i.e. systematically cheaper or more expensive configurations -
measured 12 % (mcml) to 3x (mccyl sub-grid) imbalance on 8 x B200.)
"""
import time

import numpy as np


def partition(n_configs: int, world: int, rank: int) -> np.ndarray:
    """Indices (ascending) of the configurations simulated by ``rank``: every
    ``world``-th element of a fixed permutation of the configurations."""
    n_configs, world, rank = int(n_configs), int(world), int(rank)
    if world <= 1:
        return np.arange(n_configs, dtype=np.int64)
    perm = np.random.Generator(np.random.PCG64(0x5EED0000 + n_configs)).permutation(n_configs)
    return np.sort(perm[rank::world]).astype(np.int64)


def balanced_partition(costs, world: int, rank: int) -> np.ndarray:
    """Indices (ascending) of ``rank`` for configurations of known relative cost:
    longest-processing-time-first onto the least loaded rank, ties by index - the
    same deterministic assignment on every rank; every rank gets the same number of
    configurations (+-1), as ``gather()`` expects."""
    costs = np.asarray(costs, dtype=np.float64)
    n, world = costs.size, int(world)
    if world <= 1:
        return np.arange(n, dtype=np.int64)
    order = np.lexsort((np.arange(n), -costs))           # descending cost, stable
    load = np.zeros(world)
    count = np.zeros(world, dtype=np.int64)
    cap = (n + world - 1)//world
    owner = np.empty(n, dtype=np.int64)
    for i in order:
        free = np.flatnonzero(count < cap)
        r = free[np.argmin(load[free])]
        owner[i] = r
        load[r] += costs[i]
        count[r] += 1
    return np.flatnonzero(owner == int(rank)).astype(np.int64)


class Sweep:
    def __init__(self, sim, rank: int = 0, world: int = 1):
        """``sim``: a ``pyxopto_b200`` simulator (any geometry) whose plugin
        *types* stay fixed over the sweep (what the reference requires between
        ``run`` calls as well, mc.py:486-529)."""
        self.sim = sim
        self.rank, self.world = int(rank), int(world)
        self.report = {}
        self._mine = None           # assignment of the last run (gather() reuses it)
        self.overlap = True         # two lanes in throughput mode (see run)

    def pilot_costs(self, configs, nphotons: int = 2000, apply=None) -> np.ndarray:
        """Relative cost of every configuration: loop trips of a short pilot run.
        The ranks pilot disjoint slices and exchange the counts with one all-gather
        (no exchange at world = 1).  The pilot packets are not part of any result."""
        n = len(configs)
        mine = np.arange(self.rank, n, self.world, dtype=np.int64)
        saved_rank, saved_world = self.rank, self.world
        try:
            # (run this rank's slice through the ordinary pipelined path)
            self._forced = mine
            self.run(configs, nphotons, apply=apply)
        finally:
            self._forced = None
            self.rank, self.world = saved_rank, saved_world
        local = np.zeros(n, dtype=np.float64)
        local[mine] = self.report['iterations'].astype(np.float64)
        if self.world > 1:
            import torch
            import torch.distributed as dist
            dev = torch.device('cuda', torch.cuda.current_device()) \
                if dist.get_backend() == 'nccl' else torch.device('cpu')
            t = torch.from_numpy(local).to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            local = t.cpu().numpy()
        return local

    _forced = None

    def run(self, configs, nphotons: int, apply=None, wgsize: int = None,
            maxthreads: int = None, costs=None):
        """Simulate ``nphotons`` packets for every configuration of this rank.

        ``configs``: sequence of configuration descriptors; ``apply(sim, cfg)``
        updates the simulator (layers, source, ... ) for one of them.  Without
        ``apply`` a descriptor is a dict ``{layer index: {attr: value}}``.

        Returns ``(indices, accumulators)``: the global indices of this rank's
        configurations and a ``uint64[len(indices), accumulator size]`` array of
        raw fixed-point accumulators (detector bins first, fluence after, pack
        order) - ``Sweep.detector(...)`` converts rows to reference units.

        ``costs``: relative cost per configuration (e.g. ``pilot_costs``): the
        configurations are then dealt longest-first onto the least loaded rank
        instead of by the fixed permutation."""
        from .cu import abi
        sim = self.sim
        apply = apply or _apply_layer_updates
        if self._forced is not None:
            mine = self._forced
        elif costs is not None:
            mine = balanced_partition(costs, self.world, self.rank)
        else:
            mine = partition(len(configs), self.world, self.rank)
        self._mine = (len(configs), None if costs is None else np.array(costs, copy=True))
        nphotons = int(nphotons)
        sim._ensure_device()
        # Two lanes in throughput mode: consecutive configurations run on two streams with
        # their own accumulators, counters, MWC states and packed tables (mcworker.
        # _LaneBuffers), so the CTAs of configuration k + 1 move onto the SMs that
        # configuration k has already left - a kernel of persistent threads ends with a
        # tail in which the last long-lived packets keep a few warps busy (about 1 ms of a
        # 3.5 ms kernel at 1e7 packets).  Deterministic mode keeps one lane: every
        # configuration continues the MWC states of the one before, as in the reference.
        lanes = 2 if (self.overlap and not sim.deterministic) else 1
        t0 = time.perf_counter()
        try:
            return self._run_lanes(sim, mine, configs, nphotons, apply, wgsize, maxthreads,
                                   lanes, t0)
        finally:
            sim._lane = 0
            sim._accumulator_slot = 0

    def _run_lanes(self, sim, mine, configs, nphotons, apply, wgsize, maxthreads, lanes, t0):
        from .cu import abi
        rows = counters = None
        pending = None
        slots = [None, None]
        events = [abi.Event(sim.cl_context), abi.Event(sim.cl_context)]
        for k, index in enumerate(mine):
            apply(sim, configs[int(index)])
            slot = k & 1
            if lanes == 2:
                sim._lane = slot
            else:
                sim._accumulator_slot = slot      # device double buffer
            sim.run(nphotons, wgsize=wgsize, maxthreads=maxthreads, download=False,
                    synchronize=False)
            size = int(sim.cl_rw_accumulator_allocator.size)
            if rows is None:
                rows = abi.pinned_empty(sim.cl_context, (len(mine), size), np.uint64)
                counters = abi.pinned_empty(sim.cl_context, (len(mine), 4), np.uint32)
            abuf = sim._cl_buffers[sim._rw_name('accumulator')]
            abuf.download(sim._stream, rows[k], blocking=False)
            sim._cl_buffers[sim._counters_name()].download(
                sim._stream, counters[k], blocking=False)
            events[slot].record(sim._stream)
            slots[slot] = k
            if pending is not None:
                events[pending].synchronize()     # results of configuration k-1
            pending = slot
        if pending is not None:
            events[pending].synchronize()
        for lane in range(lanes):
            sim._lane = lane
            sim._stream.synchronize()
        sim._lane = 0
        sim._accumulator_slot = 0
        iterations = counters[:, 2:4].copy().view(np.uint64).reshape(-1) \
            if rows is not None else np.zeros(0, np.uint64)
        self.report = dict(seconds=time.perf_counter() - t0, configs=len(mine),
                           packets=nphotons*len(mine), iterations=iterations,
                           threads=counters[:, 1].copy() if rows is not None else None)
        if rows is None:
            rows = np.zeros((0, 0), np.uint64)
        return mine, rows

    def detector(self, rows: np.ndarray, det, nphotons: int) -> np.ndarray:
        """Rows -> per-configuration detector data in the reference's units
        (what ``det.raw`` would hold after ``Mc.run``: weight, not normalised)."""
        sim = self.sim
        out = []
        for a in sim.cl_rw_accumulator_allocator.allocations(det):
            block = rows[:, a.offset:a.offset + a.size].astype(np.float64)
            out.append((block*(1.0/sim.types.mc_accu_k)).reshape((rows.shape[0],) + tuple(a.shape)))
        return out[0] if len(out) == 1 else out

    def gather(self, indices: np.ndarray, rows: np.ndarray, n_configs: int) -> np.ndarray:
        """All ranks get the rows of all configurations, in configuration order
        (one all-gather of the fixed-point rows over torch.distributed)."""
        if self.world == 1:
            return rows
        import torch
        import torch.distributed as dist
        per_rank = (int(n_configs) + self.world - 1)//self.world
        width = rows.shape[1]
        local = np.zeros((per_rank, width), np.int64)
        local[:rows.shape[0]] = rows.view(np.int64)
        t = torch.from_numpy(local)
        dev = torch.device('cuda', torch.cuda.current_device()) \
            if dist.get_backend() == 'nccl' else torch.device('cpu')
        t = t.to(dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        full = np.zeros((int(n_configs), width), np.uint64)
        costs = self._mine[1] if self._mine is not None and self._mine[0] == int(n_configs) \
            else None
        for r in range(self.world):
            idx = partition(n_configs, self.world, r) if costs is None \
                else balanced_partition(costs, self.world, r)
            full[idx] = out[r].cpu().numpy().view(np.uint64)[:idx.size]
        return full


def _apply_layer_updates(sim, cfg: dict):
    for layer_index, attrs in cfg.items():
        layer = sim.layers[int(layer_index)]
        for name, value in attrs.items():
            setattr(layer, name, value)


def results_from_row(sim, row: np.ndarray, nphotons: int):
    """``(None, fluence, detectors)`` result objects of one configuration from its row of
    raw fixed-point accumulators - what ``Mc.run`` returns (``mcsim._collect_results``),
    with the row in place of the per-plugin downloads.  Same conversion: ``raw =
    accumulators*(1/k)`` in float64."""
    accu_dtype = np.dtype(sim.types.np_accu)

    def blocks(obj):
        return {accu_dtype: [np.array(row[a.offset:a.offset + a.size], dtype=accu_dtype)
                             for a in sim.cl_rw_accumulator_allocator.allocations(obj)]}
    fluence_res = detectors_res = None
    if sim.fluence is not None:
        fluence_res = type(sim.fluence)(sim.fluence)
        fluence_res.update_data(sim, blocks(sim.fluence), nphotons=int(nphotons))
    if sim.detectors is not None:
        detectors_res = type(sim.detectors)(sim.detectors)
        for det, res in zip(sim.detectors, detectors_res):
            data = blocks(det)
            if data[accu_dtype]:
                detectors_res.update_data(sim, res, data, nphotons=int(nphotons))
    return None, fluence_res, detectors_res
