"""Builds, for every bench configuration, under oracle/_ref/:

  libref_<config>.so        the REFERENCE's own kernel, rendered by the reference's
                            Python and compiled unchanged with gcc -O3
                            -ffp-contract=off (IEEE evaluation, the variant the
                            parity tests compare with),
  libref_<config>_fast.so   the same text with -O3 -ffast-math -march=x86-64-v3
                            (the reference validates itself with
                            -cl-fast-relaxed-math; this is its CPU counterpart),
  inputs_<config>.npz       everything the kernel reads, produced by the
                            REFERENCE's host layer: packed plugin structs, float
                            LUT pool, MWC seeds, voxel map, buffer sizes.

TEST / BASELINE INFRASTRUCTURE; only runs where /root/reference exists.  The
files are git-ignored but travel to the GPU box, where ``bench.py --impl
reference`` and the parity tests run the reference kernel on these inputs without
importing anything of ``pyxopto_b200``.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_env  # noqa: E402

CFLAGS_IEEE = ['-O3', '-ffp-contract=off']
CFLAGS_FAST = ['-O3', '-ffast-math', '-march=x86-64-v3']
SEEDS_KEPT = 20000


def _raw(obj) -> np.ndarray:
    if obj is None:
        return np.zeros(0, np.uint8)
    if isinstance(obj, np.ndarray):
        return np.frombuffer(obj.tobytes(), np.uint8)
    return np.frombuffer(bytes(memoryview(obj).cast('B')), np.uint8)


def dump_inputs(sim, geom: str, path: str):
    """Kernel inputs of reference simulator ``sim`` as packed by the reference."""
    sizes = {}
    for n in (1000, 2000):
        sim._pack(n)
        sizes[n] = (int(sim.cl_rw_accumulator_allocator.size),
                    int(sim.cl_rw_int_allocator.size),
                    int(sim.cl_rw_float_allocator.size))
    # buffer sizes are affine in the packet count (only Trace allocates per packet)
    per_packet = [(b - a)//1000 for a, b in zip(sizes[1000], sizes[2000])]
    base = [a - 1000*k for a, k in zip(sizes[1000], per_packet)]
    sim._pack(1000)
    P = sim._packed
    out = {'packed_' + k: _raw(v) for k, v in P.items() if v is not None}
    lut = sim.float_r_lut_manager
    out['lut'] = np.ascontiguousarray(lut.pack_into(None), np.float32) if len(lut) \
        else np.zeros(4, np.float32)
    out['rng_x'] = np.ascontiguousarray(sim.rng_seeds_x[:SEEDS_KEPT])
    out['rng_a'] = np.ascontiguousarray(sim.rng_seeds_a[:SEEDS_KEPT])
    out['rmax'] = np.float32(sim.rmax)
    out['size_base'] = np.asarray(base, np.int64)
    out['size_per_packet'] = np.asarray(per_packet, np.int64)
    out['packed_at'] = np.int64(1000)
    if geom == 'mcvox':
        out['g0'] = np.uint32(len(sim.materials))
        out['voxels'] = np.ascontiguousarray(sim.voxels.data(sim)).view(np.int32)
    else:
        out['g0'] = np.uint32(len(sim.layers))
    trace = getattr(sim, 'trace', None)
    out['trace_maxlen'] = np.int64(0 if trace is None else trace.maxlen)
    np.savez_compressed(path, **out)


def main():
    if not ref_env.available():
        print('reference not present; nothing to build')
        return 0
    ref_env.activate()
    import benchcfg
    import refkernel
    for name, make in benchcfg.CONFIGS.items():
        geom = benchcfg.GEOMETRY[name]
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        sim = make(mc, cl_devices=mc.cl.Context())
        sim._pack(1000)
        sim._build_src()
        so = refkernel.compile_rendered(
            sim._cl_src, geom, name, cflags=CFLAGS_IEEE, stable=True)
        so_fast = refkernel.compile_rendered(
            sim._cl_src, geom, name + '_fast', cflags=CFLAGS_FAST, stable=True)
        dump_inputs(sim, geom, os.path.join(refkernel.REF_DIR, 'inputs_{}.npz'.format(name)))
        print('built', so, os.path.basename(so_fast), 'inputs_{}.npz'.format(name))
    return 0


if __name__ == '__main__':
    sys.exit(main())
