"""Shared helpers of the test-suite."""
import importlib
import os

import numpy as np

import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def build_sim(name: str, **kw):
    """Case ``name`` built with the pyxopto_b200 host mirror."""
    if name in cases.BENCH_RUN:
        geom, make = cases.bench_geometry(name), cases.bench_case(name)
    else:
        geom = cases.GEOMETRY.get(name) or cases.UNPINNED_GEOMETRY.get(name) or \
            cases.USER_GEOMETRY[name]
        make = cases.ALL_CASES.get(name) or cases.UNPINNED_CASES.get(name) or \
            cases.USER_CASES[name]
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geom))
    sim, attrs = make(mc, **kw)
    for k, v in attrs.items():
        setattr(sim, k, v)
    return sim, geom, mc


def run_size(name: str):
    """(packets, work-items) of the static block schedule of a case."""
    return cases.GOLDEN_RUN.get(name) or cases.UNPINNED_RUN.get(name) or \
        cases.USER_RUN.get(name) or cases.BENCH_RUN[name]


def bench_golden(name: str) -> dict:
    """Golden vectors of a bench configuration (tests/golden/bench_<name>.npz);
    the accumulator buffer is stored sparse (C3: 201^3 cells)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'bench_' + name + '.npz')))
    accu = np.zeros(int(g['accu_size']), np.uint64)
    accu[g['accu_idx']] = g['accu_val']
    g['accu'] = accu
    return g


def packed_bytes(sim) -> dict:
    out = {}
    for key, val in sim._packed.items():
        if val is None:
            continue
        out[key] = val.tobytes() if isinstance(val, np.ndarray) else \
            bytes(memoryview(val).cast('B'))
    return out


def traj_golden(name: str) -> dict:
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'traj_' + name + '.npz')))
    accu = np.zeros(int(g['accu_size']), np.uint64)
    accu[g['accu_idx']] = g['accu_val']
    g['accu'] = accu
    return g


def trajectory_agreement(rows, counts, ref_rows, ref_counts, rtol=1e-5):
    """North-star criterion "per-packet Trace trajectories within 1e-5 relative of
    the reference".  ``rows``: float32 [n, maxlen, 8] = (x, y, z, px, py, pz, w,
    pl) per event, ``counts``: events per packet.  A packet agrees when its event
    count equals the reference's and every field of every recorded event is
    within ``rtol`` of the reference, relative to the magnitude of that quantity
    over the trajectory (position: largest |coordinate|; direction: 1; weight:
    the launch weight; path length: its final value).  Returns (fraction of
    agreeing packets, largest relative deviation among them, boolean mask)."""
    n, maxlen, _ = ref_rows.shape
    rows = np.asarray(rows, np.float64).reshape(n, maxlen, 8)
    ref = np.asarray(ref_rows, np.float64)
    nrec = np.minimum(ref_counts, maxlen)
    valid = np.arange(maxlen)[None, :] < nrec[:, None]
    dev = np.abs(rows - ref)
    scale = np.ones((n, 1, 8))
    scale[:, 0, 0:3] = np.maximum(np.abs(ref[..., 0:3]).max(axis=(1, 2)), 1e-30)[:, None]
    scale[:, 0, 6] = np.maximum(np.abs(ref[..., 6]).max(axis=1), 1e-30)
    scale[:, 0, 7] = np.maximum(np.abs(ref[..., 7]).max(axis=1), 1e-30)
    rel = np.where(valid[..., None], dev/scale, 0.0).max(axis=(1, 2))
    same_count = np.asarray(counts) == np.asarray(ref_counts)
    agree = same_count & (rel <= rtol)
    worst = float(rel[agree].max()) if agree.any() else float('nan')
    return float(agree.mean()), worst, agree
