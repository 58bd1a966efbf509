"""CPU runner of the reference kernel for bench.py and the parity tests
(BASELINE / TEST INFRASTRUCTURE, not product).

Runs a bench configuration on host cores with the reference's dynamic schedule
(work-items pulling packets from an atomic counter, which is what an OpenCL CPU
runtime does with the reference kernel):
  kind "reference": oracle/_ref/libref_<config>[_fast].so, the reference's rendered
                    kernel compiled unchanged, on oracle/_ref/inputs_<config>.npz,
                    the kernel inputs packed by the reference's own host layer
                    (both built by oracle/build_ref.py).  This path imports
                    nothing of ``pyxopto_b200``: the reference process maps no
                    product library.
  kind "port":      the oracle restatement (oracle/xo_oracle.c) on structs of the
                    host mirror, only when the reference build is not available.
"""
import ctypes
import os
import time

import numpy as np

from refkernel import XoRefArgs

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')

FLAGS = {'ieee': 'gcc -O3 -ffp-contract=off',
         'fast': 'gcc -O3 -ffast-math -march=x86-64-v3'}


def _cpu_has_v3() -> bool:
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('flags'):
                    flags = set(line.split(':', 1)[1].split())
                    return {'avx2', 'fma', 'bmi2', 'movbe', 'f16c'} <= flags
    except OSError:
        pass
    return False


def variants(config: str):
    """Usable builds of the reference kernel for ``config``: [(variant, path)]."""
    out = []
    for variant, suffix in (('ieee', ''), ('fast', '_fast')):
        so = os.path.join(REF_DIR, 'libref_{}{}.so'.format(config, suffix))
        if os.path.exists(so) and (variant == 'ieee' or _cpu_has_v3()):
            out.append((variant, so))
    return out


def available(config: str) -> bool:
    return bool(variants(config)) and \
        os.path.exists(os.path.join(REF_DIR, 'inputs_{}.npz'.format(config)))


_inputs_cache = {}


def load_inputs(config: str) -> dict:
    if config not in _inputs_cache:
        with np.load(os.path.join(REF_DIR, 'inputs_{}.npz'.format(config))) as f:
            _inputs_cache[config] = {k: f[k] for k in f.files}
    return _inputs_cache[config]


def run_reference(config: str, geometry: str, nphotons: int, threads: int,
                  variant: str = 'ieee', seed_offset: int = 0):
    """The reference kernel on the reference-packed inputs.  Returns
    dict(accu, ints, floats, seconds, items)."""
    nphotons = int(nphotons)
    so = dict(variants(config))[variant]
    lib = ctypes.CDLL(so)
    lib.xo_ref_run_dynamic_items.argtypes = [ctypes.POINTER(XoRefArgs), ctypes.c_uint32,
                                             ctypes.c_uint32]
    lib.xo_ref_run_dynamic_items.restype = None
    inp = load_inputs(config)
    nseeds = len(inp['rng_x']) - int(seed_offset)
    # work-items: one per host thread; mccyl gives each work-item a budget of 1e6
    # loop trips (mccyl.template.c:681), so it needs many more
    nitems = int(threads)
    if geometry == 'mccyl':
        nitems = max(threads, nphotons//1000 + threads)
    nitems = int(min(nitems, nseeds))
    threads = int(min(threads, nitems))
    keep = []

    def buf(key, minsize=16):
        raw = inp[key].tobytes() if key in inp else b''
        b = ctypes.create_string_buffer(raw if raw else b'\0'*minsize, max(len(raw), minsize))
        keep.append(b)
        return ctypes.addressof(b)

    sizes = inp['size_base'] + inp['size_per_packet']*nphotons
    accu = np.zeros(max(int(sizes[0]), 1), np.uint64)
    ints = np.zeros(max(int(sizes[1]), 1), np.int32)
    floats = np.zeros(max(int(sizes[2]), 1), np.float32)
    lut = np.ascontiguousarray(inp['lut'], np.float32)
    done = np.zeros(1, np.uint32)
    nk = np.zeros(1, np.uint32)
    x = inp['rng_x'][seed_offset:seed_offset + nitems].copy()
    a = np.ascontiguousarray(inp['rng_a'][seed_offset:seed_offset + nitems])
    args = XoRefArgs()
    args.num_packets = nphotons
    args.num_packets_done = done.ctypes.data
    args.num_kernels = nk.ctypes.data
    args.rmax = float(inp['rmax'])
    args.rng_x, args.rng_a = x.ctypes.data, a.ctypes.data
    args.g0 = int(inp['g0'])
    if geometry == 'mcvox':
        vox = np.ascontiguousarray(inp['voxels'])
        keep.append(vox)
        args.g1 = buf('packed_voxels')
        args.g2 = vox.ctypes.data
        args.g3 = buf('packed_materials')
    else:
        args.g1 = buf('packed_layers')
    args.source = buf('packed_source')
    args.surface = buf('packed_surface_layouts')
    args.trace = buf('packed_trace')
    args.fluence = buf('packed_fluence')
    args.detectors = buf('packed_detectors')
    args.fp_lut = lut.ctypes.data
    args.int_buffer = ints.ctypes.data
    args.float_buffer = floats.ctypes.data
    args.accumulator_buffer = accu.ctypes.data
    t = time.perf_counter()
    lib.xo_ref_run_dynamic_items(ctypes.byref(args), threads, nitems)
    dt = time.perf_counter() - t
    assert int(done[0]) >= nphotons, (int(done[0]), nphotons)
    return dict(accu=accu, ints=ints, floats=floats, seconds=dt, items=nitems)


def run(config: str, geometry: str, nphotons: int, threads: int, variant: str = None):
    """Times ``nphotons`` packets of ``config`` on ``threads`` host threads.
    Returns (packets/s, kind, seconds, flags).  ``variant`` None: the IEEE build."""
    nphotons = int(nphotons)
    if available(config):
        usable = dict(variants(config))
        variant = variant if variant in usable else 'ieee'
        res = run_reference(config, geometry, nphotons, threads, variant)
        return nphotons/res['seconds'], 'reference', res['seconds'], FLAGS[variant]
    # the reference build is absent: oracle restatement on the host mirror's structs
    import importlib
    import benchcfg
    import xo_oracle
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geometry))
    sim = benchcfg.CONFIGS[config](mc)
    sim._pack(nphotons)
    desc = xo_oracle.describe(sim, geometry)
    t = time.perf_counter()
    xo_oracle.run(desc, nphotons, threads, sim.rng_seeds_x[:threads],
                  sim.rng_seeds_a[:threads], math=xo_oracle.MATH_LIBM,
                  schedule='dynamic')
    dt = time.perf_counter() - t
    return nphotons/dt, 'port', dt, 'gcc -O2 (oracle port)'


def fastest_variant(config: str, geometry: str, pilot: int, threads: int):
    """(variant, packets/s) of the faster usable build on a pilot sample."""
    best = (None, 0.0)
    if not available(config):
        return best
    for variant, _ in variants(config):
        pps = run(config, geometry, pilot, threads, variant)[0]
        if pps > best[1]:
            best = (variant, pps)
    return best
