"""CPU baseline runner for bench.py (BASELINE INFRASTRUCTURE, not product).

Runs the bench workload on host cores with the reference's dynamic schedule
(one work-item per host thread pulling packets from an atomic counter, which is
what an OpenCL CPU runtime does with the reference kernel):
  kind "reference": oracle/_ref/libref_<config>.so, the reference's rendered
                    kernel compiled unchanged (built by oracle/build_ref.py);
  kind "port":      the oracle restatement (oracle/xo_oracle.c), when the
                    reference build is not available.
The packed plugin structs come from the pyxopto_b200 host mirror, whose bytes
are pinned to the reference's by tests/test_host_packing.py.
"""
import ctypes
import os
import time

import numpy as np

import xo_oracle
from refkernel import XoRefArgs, GEOMETRY_ID

HERE = os.path.dirname(os.path.abspath(__file__))


def _raw(obj):
    return bytes(memoryview(obj).cast('B')) if obj is not None else b''


def run(sim, geometry: str, config: str, nphotons: int, threads: int):
    """Returns (packets/s, kind, seconds)."""
    nphotons = int(nphotons)
    sim._pack(nphotons)
    so = os.path.join(HERE, '_ref', 'libref_{}.so'.format(config))
    lib = None
    if os.path.exists(so):
        try:
            lib = ctypes.CDLL(so)
        except OSError:          # unusable build: fall back to the port
            lib = None
    if lib is not None:
        lib.xo_ref_run_dynamic_items.argtypes = [ctypes.POINTER(XoRefArgs), ctypes.c_uint32,
                                                 ctypes.c_uint32]
        lib.xo_ref_run_dynamic_items.restype = None
        # work-items: one per host thread; mccyl gives each work-item a budget of
        # 1e6 loop trips (mccyl.template.c:681), so it needs many more
        nitems = int(threads)
        if geometry == 'mccyl':
            nitems = int(min(max(threads, nphotons//1000 + threads), len(sim.rng_seeds_x)))
        keep = []

        def buf(raw, minsize=16):
            b = ctypes.create_string_buffer(raw if raw else b'\0'*minsize, max(len(raw), minsize))
            keep.append(b)
            return ctypes.addressof(b)

        P = sim._packed
        accu = np.zeros(max(int(sim.cl_rw_accumulator_allocator.size), 1), np.uint64)
        ints = np.zeros(max(int(sim.cl_rw_int_allocator.size), 1), np.int32)
        floats = np.zeros(max(int(sim.cl_rw_float_allocator.size), 1), np.float32)
        lut = sim.float_r_lut_manager.pack_into(None).astype(np.float32)
        done = np.zeros(1, np.uint32)
        nk = np.zeros(1, np.uint32)
        x = sim.rng_seeds_x[:nitems].copy()
        a = np.ascontiguousarray(sim.rng_seeds_a[:nitems])
        args = XoRefArgs()
        args.num_packets = nphotons
        args.num_packets_done = done.ctypes.data
        args.num_kernels = nk.ctypes.data
        args.rmax = np.float32(sim.rmax)
        args.rng_x, args.rng_a = x.ctypes.data, a.ctypes.data
        if geometry == 'mcvox':
            vox = np.ascontiguousarray(sim.voxels.data(sim))
            keep.append(vox)
            args.g0 = len(sim.materials)
            args.g1 = buf(_raw(P['voxels']))
            args.g2 = vox.ctypes.data
            args.g3 = buf(_raw(P['materials']))
        else:
            args.g0 = len(sim.layers)
            args.g1 = buf(_raw(P['layers']))
        args.source = buf(_raw(P['source']))
        args.surface = buf(b'')
        args.trace = buf(_raw(P.get('trace')))
        args.fluence = buf(_raw(P.get('fluence')))
        args.detectors = buf(_raw(P.get('detectors')))
        args.fp_lut = lut.ctypes.data
        args.int_buffer = ints.ctypes.data
        args.float_buffer = floats.ctypes.data
        args.accumulator_buffer = accu.ctypes.data
        t = time.perf_counter()
        lib.xo_ref_run_dynamic_items(ctypes.byref(args), int(threads), nitems)
        dt = time.perf_counter() - t
        assert int(done[0]) >= nphotons, (int(done[0]), nphotons)
        return nphotons/dt, 'reference', dt
    desc = xo_oracle.describe(sim, geometry)
    t = time.perf_counter()
    xo_oracle.run(desc, nphotons, threads, sim.rng_seeds_x[:threads],
                  sim.rng_seeds_a[:threads], math=xo_oracle.MATH_LIBM,
                  schedule='dynamic')
    dt = time.perf_counter() - t
    return nphotons/dt, 'port', dt
