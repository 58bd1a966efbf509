"""Genuine ``xopto`` plugin objects handed to this engine (runs where the reference
is installed, /root/reference; skipped on the GPU box).

Every case of tests/cases.py is built with the REFERENCE package; the plugin
objects of that simulator (layers / materials / voxels, source, detectors, fluence,
trace, surface layouts, options) then go, as they are, into the constructor of the
matching ``pyxopto_b200`` simulator.  The engine rebuilds them from their own
``todict()`` description (pyxopto_b200/adopt.py); the packed structs must equal the
reference's own packing byte for byte (= the golden vectors) and the CUDA
translation unit must compile for sm_100a."""
import importlib
import os
import sys

import numpy as np
import pytest

import cases
from helpers import golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                'oracle'))
import ref_env  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_env.available(),
                                reason='needs the reference package (/root/reference)')

def _same_struct(mine: bytes, ref: bytes) -> bool:
    """Byte-identical, or - for the direction vectors the reference's own
    serialisation already normalised once (todict() stores the unit vector, the
    constructor normalises again) - fp32 fields within 2 ulp."""
    if mine == ref:
        return True
    if len(mine) != len(ref) or len(ref) % 4:
        return False
    a, b = np.frombuffer(mine, np.uint32), np.frombuffer(ref, np.uint32)
    diff = np.flatnonzero(a != b)
    fa, fb = a.view(np.float32)[diff], b.view(np.float32)[diff]
    return bool(np.all(np.isfinite(fa)) and np.all(np.isfinite(fb)) and
                np.all(np.abs(a[diff].astype(np.int64) - b[diff].astype(np.int64)) <= 2))


NAMES = sorted(n for n in cases.ALL_CASES if n not in getattr(cases, 'USER_CASES', {}) and
               not n.startswith('mcml_user'))


def _reference_sim(name):
    ref_env.activate()
    geom = cases.GEOMETRY[name]
    rmc = importlib.import_module('xopto.{}.mc'.format(geom))
    sim, attrs = cases.ALL_CASES[name](rmc, cl_devices=rmc.cl.Context())
    return sim, attrs, geom


@pytest.mark.parametrize('name', NAMES)
def test_reference_objects_pack_and_compile(name):
    ref, attrs, geom = _reference_sim(name)
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geom))
    kw = dict(detectors=ref.detectors, trace=ref.trace, fluence=ref.fluence,
              options=list(getattr(ref, '_options', None) or []),
              rnginit=int(ref.rng_seeds_x[0]) if False else None)
    if geom == 'mcml':
        sim = mc.Mc(ref.layers, ref.source, surface=ref.surface, **kw)
    elif geom == 'mccyl':
        sim = mc.Mc(ref.layers, ref.source, **kw)
    else:
        sim = mc.Mc(ref.voxels, ref.materials, ref.source, **kw)
    for k, v in attrs.items():
        setattr(sim, k, v)
    for obj in (sim.source, sim.detectors, sim.trace, sim.fluence):
        assert obj is None or type(obj).__module__.startswith('pyxopto_b200')
    g = golden(name)
    n = int(g['nphotons'])
    sim._pack(n)
    from pyxopto_b200.cl import cltypes
    for key in [k for k in g.files if k.startswith('packed_')]:
        mine = sim._packed.get(key[len('packed_'):])
        assert mine is not None, key
        assert _same_struct(cltypes.raw_bytes(mine), g[key].tobytes()), key
    if len(sim._float_lut):
        lut = sim._float_lut.pack_into(None).astype(np.float32)
        assert np.array_equal(lut[:g['lut'].size], g['lut'][:lut.size])
    cubin, _, _ = sim.compile(n, block=64)
    assert len(cubin) > 0
