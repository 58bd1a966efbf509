"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``pyxopto_b200`` never does.

The oracle consumes the *packed plugin structs* (raw bytes, reference ctypes
layout) of a simulator object and maps the plugin classes - by class name, so
the mapping works for reference objects and for the pyxopto_b200 host mirror
alike - onto the run-time "kind" switches of ``xo_oracle.c``.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libxo_oracle.so')
SOURCES = ['xo_oracle.c', 'xo_oracle.h', 'xo_detmath.h',
           'xo_oracle_vox.inc', 'xo_oracle_cyl.inc', 'xo_oracle_sv.inc']

GEOMETRY = {'mcml': 0, 'mcvox': 1, 'mccyl': 2}
METHOD = {'aw': 0, 'albedo_weight': 0, 'ar': 1, 'albedo_rejection': 1,
          'mbl': 2, 'microscopic_beer_lambert': 2}
MATH_LIBM, MATH_PORTABLE = 0, 1
PF_KIND = {'Hg': 1, 'MHg': 2, 'Gk': 3, 'Lut': 4, 'LutEx': 4, 'Hg2': 5, 'Gk2': 6,
           'MGk': 7, 'Pc': 8, 'MPc': 9, 'HgDir': 10, 'Rayleigh': 11}
SRC_KIND = {'Line': 1, 'GaussianBeam': 2, 'UniformFiber': 3,
            'IsotropicPoint': 4, 'UniformBeam': 5, 'LambertianFiber': 6,
            'IsotropicVoxel': 7, 'UniformFiberLut': 8,
            'UniformRectangular': 9, 'LambertianRectangular': 10,
            'IsotropicVoxels': 11, 'UniformFiberNI': 12, 'LambertianFiberNI': 13,
            'UniformFiberLutNI': 14, 'UniformRectangularLut': 15}
DET_KIND = {'NoneType': 0, 'DetectorDefault': 0, 'Total': 1, 'Radial': 2,
            'Cartesian': 3, 'SixAroundOne': 4, 'RadialPl': 5, 'TotalPl': 6,
            'SymmetricX': 7, 'FiZ': 8, 'CartesianPl': 9, 'SixAroundOnePl': 10,
            'LinearArray': 12, 'FiberArray': 13, 'LinearArrayPl': 14, 'FiberArrayPl': 15, 'TotalLut': 16, 'TotalLutPl': 17,
            'FiberLutArray': 18}
DET_KIND_TOTAL_CYL = 11
SURF_KIND = {'NoneType': 0, 'SurfaceLayoutDefault': 0, 'LambertianReflector': 1,
             'SixAroundOne': 2, 'LinearArray': 3, 'FiberArray': 4}
FLU_KIND = {'NoneType': 0, 'Fluence': 1, 'FluenceRz': 2, 'Fluencet': 3,
            'FluenceRzt': 4, 'FluenceCyl': 5, 'FluenceCylt': 6}


class Job(ctypes.Structure):
    _fields_ = [
        ('geometry', ctypes.c_int32), ('method', ctypes.c_int32),
        ('math', ctypes.c_int32), ('use_lottery', ctypes.c_int32),
        ('weight_min', ctypes.c_float), ('lottery_chance', ctypes.c_float),
        ('pf_kind', ctypes.c_int32), ('pf_size', ctypes.c_int32),
        ('src_kind', ctypes.c_int32),
        ('det_kind', ctypes.c_int32*3), ('det_offset', ctypes.c_int32*3),
        ('det_param', ctypes.c_int32*3),
        ('fluence_kind', ctypes.c_int32), ('fluence_rate', ctypes.c_int32),
        ('trace_flags', ctypes.c_int32), ('use_events', ctypes.c_int32),
        ('track_opl', ctypes.c_int32),
        ('surf_kind', ctypes.c_int32*2), ('surf_offset', ctypes.c_int32*2),
        ('surf_param', ctypes.c_int32*2), ('enhanced_rng', ctypes.c_int32),
        ('anisotropic', ctypes.c_int32),
        ('num_packets', ctypes.c_uint32), ('num_threads', ctypes.c_uint32),
        ('rmax', ctypes.c_float), ('num_layers', ctypes.c_uint32),
        ('layers', ctypes.c_void_p), ('voxel_cfg', ctypes.c_void_p),
        ('voxels', ctypes.c_void_p), ('source', ctypes.c_void_p),
        ('detectors', ctypes.c_void_p), ('fluence', ctypes.c_void_p),
        ('trace', ctypes.c_void_p), ('surface', ctypes.c_void_p),
        ('fp_lut', ctypes.c_void_p),
        ('rng_x', ctypes.c_void_p), ('rng_a', ctypes.c_void_p),
        ('int_buffer', ctypes.c_void_p), ('float_buffer', ctypes.c_void_p),
        ('accumulator_buffer', ctypes.c_void_p),
        ('num_kernels', ctypes.c_uint32), ('num_packets_done', ctypes.c_uint32),
        ('num_iterations', ctypes.c_uint64),
    ]


def build(force: bool = False, fast: bool = False) -> str:
    """Compile the oracle with gcc.  ``fast`` builds the -O3 -ffast-math
    variant used only as a CPU *baseline* (never as a checker)."""
    out = LIB_PATH if not fast else LIB_PATH.replace('.so', '_fast.so')
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    if not force and os.path.exists(out) and all(
            os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    flags = (['-O3', '-march=native', '-ffast-math'] if fast
             else ['-O2', '-ffp-contract=off'])
    cmd = ['gcc', '-std=gnu11', '-fPIC', '-shared', '-pthread'] + flags + \
          [srcs[0], '-o', out + '.tmp', '-lm']
    subprocess.check_call(cmd)
    os.replace(out + '.tmp', out)
    return out


_libs = {}


def lib(fast: bool = False):
    if fast not in _libs:
        L = ctypes.CDLL(build(fast=fast))
        L.xo_oracle_run.argtypes = [ctypes.POINTER(Job)]
        L.xo_oracle_run.restype = ctypes.c_int
        L.xo_oracle_run_dynamic.argtypes = [ctypes.POINTER(Job), ctypes.c_uint32]
        L.xo_oracle_run_dynamic.restype = ctypes.c_int
        L.xo_oracle_rng_test.argtypes = [
            ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.xo_oracle_rng_test.restype = None
        L.xo_oracle_init_rng.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint32, ctypes.c_uint64]
        L.xo_oracle_init_rng.restype = ctypes.c_int
        L.xo_oracle_math_probe.argtypes = [
            ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32] + [ctypes.c_void_p]*4
        L.xo_oracle_math_probe.restype = None
        L.xo_oracle_sampling_volume.argtypes = [ctypes.c_uint32] + [ctypes.c_void_p]*6
        L.xo_oracle_sampling_volume.restype = ctypes.c_uint64
        _libs[fast] = L
    return _libs[fast]


def rng_test(x: int, a: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    lib().xo_oracle_rng_test(int(x), int(a), n, out.ctypes.data)
    return out


def init_rng(fora: np.ndarray, n: int, xinit: int):
    fora = np.ascontiguousarray(fora, dtype=np.uint32)
    x = np.zeros(n, np.uint64)
    a = np.zeros(n, np.uint32)
    rc = lib().xo_oracle_init_rng(x.ctypes.data, a.ctypes.data, fora.ctypes.data,
                                  n, int(xinit))
    if rc:
        raise ValueError('invalid xinit')
    return x, a


MATH_FN = {'log': 0, 'sincos': 1, 'cbrt': 2, 'pow': 3, 'exp': 4, 'atan2': 5,
           'sqrt': 6, 'div': 7}


def math_probe(fn: str, math: int, in0, in1=None):
    in0 = np.ascontiguousarray(in0, np.float32)
    in1 = np.ascontiguousarray(in0 if in1 is None else in1, np.float32)
    out0 = np.zeros_like(in0)
    out1 = np.zeros_like(in0)
    lib().xo_oracle_math_probe(MATH_FN[fn], math, in0.size, in0.ctypes.data,
                               in1.ctypes.data, out0.ctypes.data, out1.ctypes.data)
    return out0, out1


def _raw(obj) -> bytes:
    if obj is None:
        return b''
    if isinstance(obj, (bytes, bytearray)):
        return bytes(obj)
    if isinstance(obj, np.ndarray):
        return obj.tobytes()
    return bytes(memoryview(obj).cast('B'))


def _name(obj) -> str:
    return type(obj).__name__


def describe(mc_obj, geometry: str) -> dict:
    """Plugin kinds + packed bytes of a simulator object whose ``_pack`` has
    been called (reference ``Mc`` or pyxopto_b200 ``Mc``)."""
    P = mc_obj._packed
    if geometry == 'mcvox':
        media = mc_obj.materials
        pf = media[0].pf
        layers_key, n_media = 'materials', len(media)
    else:
        media = mc_obj.layers
        pf = media[1].pf
        layers_key, n_media = 'layers', len(media)
    pf_type = pf.fetch_cl_type(mc_obj) if hasattr(pf, 'fetch_cl_type') else pf.cl_type(mc_obj)
    aniso = int(_name(media[1 if geometry != 'mcvox' else 0]).startswith('Anisotropic'))
    d = dict(anisotropic=aniso, geometry=geometry, pf_kind=PF_KIND[_name(pf)],
             pf_size=ctypes.sizeof(pf_type), num_layers=n_media,
             layers=_raw(P[layers_key]), source=_raw(P['source']),
             src_kind=SRC_KIND[_name(mc_obj.source)])
    dets = mc_obj.detectors
    det_kind, det_off, det_par = [0, 0, 0], [0, 0, 0], [0, 0, 0]
    if dets is not None:
        dstruct = type(P['detectors'])
        # mccyl: `outer` takes the slot of `top`, there is no `bottom`
        locs = ('outer', None, 'specular') if geometry == 'mccyl' \
            else ('top', 'bottom', 'specular')
        for i, loc in enumerate(locs):
            if loc is None:
                continue
            det = getattr(dets, loc)
            det_kind[i] = DET_KIND[_name(det)]
            if geometry == 'mccyl' and det_kind[i] == DET_KIND['Total']:
                det_kind[i] = DET_KIND_TOTAL_CYL
            det_off[i] = getattr(dstruct, loc).offset
            det_par[i] = int(getattr(det, 'n', 0) or 0) if det_kind[i] in (12, 13, 14, 15, 18) else 0
        d['detectors'] = _raw(P['detectors'])
    d['det_kind'], d['det_offset'], d['det_param'] = det_kind, det_off, det_par
    if hasattr(mc_obj, 'resolved_options'):
        d['enhanced_rng'] = int(bool(mc_obj.resolved_options().get('MC_USE_ENHANCED_RNG', False)))
    surf = getattr(mc_obj, 'surface', None)
    d['surf_kind'], d['surf_offset'] = [0, 0], [0, 0]
    if surf is not None and geometry == 'mcml':
        packed = P.get('surface_layouts', P.get('surface'))     # reference / this repo
        sstruct = type(packed)
        for i, loc in enumerate(('top', 'bottom')):
            d['surf_kind'][i] = SURF_KIND[_name(getattr(surf, loc))]
            d.setdefault('surf_param', [0, 0])[i] = int(getattr(getattr(surf, loc), 'n', 0) or 0) \
                if d['surf_kind'][i] in (3, 4) else 0
            d['surf_offset'][i] = getattr(sstruct, loc).offset
        d['surface'] = _raw(packed)
    flu = mc_obj.fluence
    d['fluence_kind'] = FLU_KIND[_name(flu)]
    if flu is not None:
        d['fluence'] = _raw(P['fluence'])
        d['fluence_rate'] = int(flu.mode == 'fluence')
    tr = mc_obj.trace
    track_opl = any(k in (5, 6, 9, 10, 14, 15, 17) for k in det_kind) or \
        d['fluence_kind'] in (3, 4, 6)
    if tr is not None:
        d['trace'] = _raw(P['trace'])
        d['trace_flags'] = int(tr.options)
        d['use_events'] = int(tr.event_mask is not None)
        d['trace_maxlen'] = int(tr.maxlen)
        track_opl = track_opl or bool(tr.plon)
    d['track_opl'] = int(track_opl)
    if geometry == 'mcvox':
        d['voxel_cfg'] = _raw(P['voxels'])
        d['voxels'] = np.ascontiguousarray(mc_obj.voxels.data(mc_obj)).view(np.int32)
    d['rmax'] = float(mc_obj.rmax)
    d['sizes'] = (int(mc_obj.cl_rw_accumulator_allocator.size),
                  int(mc_obj.cl_rw_int_allocator.size),
                  int(mc_obj.cl_rw_float_allocator.size))
    mgr = mc_obj.float_r_lut_manager
    d['fp_lut'] = (np.ascontiguousarray(mgr.pack_into(None), dtype=np.float32)
                   if len(mgr) else np.zeros(1, np.float32))
    return d


def run(desc: dict, nphotons: int, nthreads: int, rng_x: np.ndarray,
        rng_a: np.ndarray, math: int = MATH_LIBM, method: str = 'aw',
        use_lottery: bool = True, weight_min: float = 1e-4,
        lottery_chance: float = 0.1, schedule: str = 'static',
        fast: bool = False) -> dict:
    """Run the oracle; returns dict(accu, ints, floats, rng_x, num_kernels,
    done, iterations)."""
    job = Job()
    keep = []

    def buf(raw: bytes):
        b = ctypes.create_string_buffer(raw if raw else b'\0'*16, max(len(raw), 16))
        keep.append(b)
        return ctypes.addressof(b)

    job.geometry = GEOMETRY[desc['geometry']]
    job.method = METHOD[method]
    job.math = math
    job.use_lottery = int(use_lottery)
    job.weight_min = weight_min
    job.lottery_chance = lottery_chance
    job.pf_kind, job.pf_size = desc['pf_kind'], desc['pf_size']
    job.src_kind = desc['src_kind']
    for i in range(3):
        job.det_kind[i] = desc['det_kind'][i]
        job.det_offset[i] = desc['det_offset'][i]
        job.det_param[i] = desc.get('det_param', [0, 0, 0])[i]
    job.fluence_kind = desc.get('fluence_kind', 0)
    job.fluence_rate = desc.get('fluence_rate', 0)
    job.trace_flags = desc.get('trace_flags', 0)
    job.use_events = desc.get('use_events', 0)
    job.track_opl = desc.get('track_opl', 0)
    job.enhanced_rng = int(desc.get('enhanced_rng', 0))
    job.anisotropic = int(desc.get('anisotropic', 0))
    job.num_packets = int(nphotons)
    job.num_threads = int(nthreads)
    job.rmax = np.float32(desc['rmax'])
    job.num_layers = desc['num_layers']
    job.layers = buf(desc['layers'])
    job.source = buf(desc['source'])
    job.detectors = buf(desc.get('detectors', b''))
    job.fluence = buf(desc.get('fluence', b''))
    job.trace = buf(desc.get('trace', b''))
    job.surface = buf(desc.get('surface', b''))
    for i in range(2):
        job.surf_kind[i] = desc.get('surf_kind', [0, 0])[i]
        job.surf_offset[i] = desc.get('surf_offset', [0, 0])[i]
        job.surf_param[i] = desc.get('surf_param', [0, 0])[i]
    if desc['geometry'] == 'mcvox':
        job.voxel_cfg = buf(desc['voxel_cfg'])
        vox = np.ascontiguousarray(desc['voxels'], np.int32)
        keep.append(vox)
        job.voxels = vox.ctypes.data
    lut = np.ascontiguousarray(desc['fp_lut'], np.float32)
    job.fp_lut = lut.ctypes.data
    na, ni, nf = desc['sizes']
    accu = np.zeros(max(na, 1), np.uint64)
    ints = np.zeros(max(ni, 1), np.int32)
    floats = np.zeros(max(nf, 1), np.float32)
    x = np.array(rng_x, dtype=np.uint64, copy=True)
    a = np.ascontiguousarray(rng_a, dtype=np.uint32)
    job.rng_x, job.rng_a = x.ctypes.data, a.ctypes.data
    job.int_buffer = ints.ctypes.data
    job.float_buffer = floats.ctypes.data
    job.accumulator_buffer = accu.ctypes.data
    L = lib(fast)
    if schedule == 'static':
        rc = L.xo_oracle_run(ctypes.byref(job))
    else:
        rc = L.xo_oracle_run_dynamic(ctypes.byref(job), int(nthreads))
    if rc:
        raise RuntimeError('oracle failed with code %d' % rc)
    return dict(accu=accu, ints=ints, floats=floats, rng_x=x,
                num_kernels=int(job.num_kernels), done=int(job.num_packets_done),
                iterations=int(job.num_iterations))


def sampling_volume(trace_packed, sv_packed, npackets: int, ints: np.ndarray,
                    floats: np.ndarray, accu_size: int) -> dict:
    """SamplingVolume kernel restated: returns dict(accu, total_weight, steps)."""
    tp, sp = _raw(trace_packed), _raw(sv_packed)
    tb = ctypes.create_string_buffer(tp, len(tp))
    sb = ctypes.create_string_buffer(sp, len(sp))
    ints = np.ascontiguousarray(ints, np.int32)
    floats = np.ascontiguousarray(floats, np.float32)
    accu = np.zeros(max(int(accu_size), 1), np.uint64)
    total = np.zeros(1, np.uint64)
    steps = lib().xo_oracle_sampling_volume(
        int(npackets), ctypes.addressof(tb), ctypes.addressof(sb), total.ctypes.data,
        ints.ctypes.data, floats.ctypes.data, accu.ctypes.data)
    return dict(accu=accu, total_weight=int(total[0]), steps=int(steps))
