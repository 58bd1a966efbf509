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
    geom = cases.GEOMETRY.get(name) or cases.UNPINNED_GEOMETRY.get(name) or \
        cases.USER_GEOMETRY[name]
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geom))
    make = cases.ALL_CASES.get(name) or cases.UNPINNED_CASES.get(name) or \
        cases.USER_CASES[name]
    sim, attrs = make(mc, **kw)
    for k, v in attrs.items():
        setattr(sim, k, v)
    return sim, geom, mc


def run_size(name: str):
    """(packets, work-items) of the static block schedule of a case."""
    return cases.GOLDEN_RUN.get(name) or cases.UNPINNED_RUN.get(name) or cases.USER_RUN[name]


def packed_bytes(sim) -> dict:
    out = {}
    for key, val in sim._packed.items():
        if val is None:
            continue
        out[key] = val.tobytes() if isinstance(val, np.ndarray) else \
            bytes(memoryview(val).cast('B'))
    return out
