"""Host-side decisions of the mcvox throughput path (no GPU): which photon loop a
configuration compiles - the packet pool (mcvox_pool_loop.cuh), the lane-resident rays
(mcvox_dda_loop.cuh) or the reference-structured loop - and the shared memory it asks for."""
import numpy as np
import pytest

import benchcfg
import cases
from helpers import build_sim
from pyxopto_b200.mcbase import mcoptions
from pyxopto_b200.mcvox import mc as voxmc


def _define(src: str, name: str) -> str:
    for line in src.splitlines():
        if line.startswith('#define {} '.format(name)):
            return line.split()[2]
    raise KeyError(name)


def test_c3_compiles_the_packet_pool():
    sim = benchcfg.c3_vox(voxmc, n=21)
    sim._pack(1000)
    assert sim._loop_name() == 'packet pool'
    src = sim.kernel_source(block=1024)
    assert _define(src, 'XO_VOX_POOL') == '64'
    assert _define(src, 'XO_USE_RMAX') == '0'
    # per warp: 64 slots of 17 words + 1 state byte, 32 bytes of gather indices, five rings of
    # 64 slot numbers
    assert sim._queue_bytes(1024) == 32*(64*69 + 32 + 320) + 32
    assert sim._queue_bytes(1024) + sim._shared_layout(sim._medium_bytes())[0] < 227*1024


@pytest.mark.parametrize('change, loop', [
    (lambda sim: setattr(sim, 'pool_slots', 0), 'lane-resident rays'),
    (lambda sim: setattr(sim, 'rmax', 50e-6), 'packet pool'),       # sphere inside the box
    (lambda sim: sim._options.append(mcoptions.McMethod.mbl), 'reference-structured'),
    (lambda sim: sim._options.append(mcoptions.McDeterministic.on), 'reference-structured'),
    (lambda sim: sim._options.append(mcoptions.McMethod.ar), 'packet pool'),
])
def test_loop_selection(change, loop):
    sim = benchcfg.c3_vox(voxmc, n=21)
    change(sim)
    sim._pack(1000)
    assert sim._loop_name() == loop
    src = sim.kernel_source(block=256)
    assert (_define(src, 'XO_VOX_POOL') == '64') == (loop == 'packet pool')
    if loop != 'packet pool':
        assert sim._queue_bytes(256) == 40*256 + 16
    else:
        # (+ one float per slot where the rmax sphere can be reached)
        rmax = int(_define(src, 'XO_USE_RMAX'))
        assert sim._queue_bytes(256) == 8*(64*(69 + 4*rmax) + 32 + 320) + 32


def test_traces_and_the_pool():
    """A full trace keeps the lane-resident loop (store-bound: measured slower in the pool)
    unless asked for; the pool then carries a trace quad per slot."""
    sim, geom, _ = build_sim('mcvox_line_mhg_trace')
    sim._pack(600)
    assert geom == 'mcvox' and sim._loop_name() == 'lane-resident rays'
    sim.pool_full_trace = True
    assert sim._loop_name() == 'packet pool'
    assert _define(sim.kernel_source(block=64), 'XO_VOX_POOL') == '64'
    # per warp: 64 slots of 5 quads + 1 state byte, 32 bytes of gather indices
    assert sim._queue_bytes(64) == 2*(64*81 + 32 + 320) + 32
    sim.pool_slots = 0
    assert sim._loop_name() == 'lane-resident rays' and sim._queue_bytes(64) == 40*64 + 16


def test_pool_thresholds_reach_the_translation_unit():
    sim = benchcfg.c3_vox(voxmc, n=21)
    sim.pool_tuning = {'XO_POOL_THR_W': 10}
    sim._pack(1000)
    assert _define(sim.kernel_source(block=1024), 'XO_POOL_THR_W') == '10'


def test_fluence_result_materializes_a_pending_device_grid_on_access():
    """``Mc.lazy_fluence`` hook of the result objects: whatever touches ``_data`` (``raw``,
    ``data``, the setters, the copy constructor, ``update``) first runs the pending loader."""
    from pyxopto_b200.mcbase import mcfluence
    from pyxopto_b200.mcbase.mcutil.axis import Axis
    calls = []

    def loader(f):
        calls.append(f)
        f._store = np.full(f.shape, 2.0)

    flu = mcfluence.Fluence(Axis(0, 1, 2), Axis(0, 1, 3), Axis(0, 1, 4))
    assert flu.raw is None and flu._pending is None
    flu._pending = loader
    assert flu.raw.shape == (4, 3, 2) and calls == [flu] and flu._pending is None
    assert flu.raw is flu.raw and len(calls) == 1
    # a setter collects the pending grid before it overwrites it
    flu._pending = loader
    flu.raw = np.zeros(flu.shape)
    assert len(calls) == 2 and float(flu.raw.sum()) == 0.0
    # copies and updates see the collected data
    flu._pending = loader
    cp = mcfluence.Fluence(flu)
    assert len(calls) == 3 and float(cp.raw.sum()) == 2.0*24 and cp._pending is None
    flu._pending = loader
    cp.update(flu)
    assert len(calls) == 4 and float(cp.raw.sum()) == 4.0*24
    rz = mcfluence.FluenceRz(Axis(0, 1, 5), Axis(0, 1, 6))
    rz._pending = lambda f: setattr(f, '_store', np.ones(30))
    assert rz.raw_zr.shape == (6, 5) and rz._pending is None
