"""Run the *reference's own kernel* on CPU cores (TEST INFRASTRUCTURE).

``RefKernel(mc_obj)`` takes an ``Mc`` instance created with the reference package
(see ``ref_env.activate``), lets the reference pack its structs and render its
OpenCL-C translation unit (``mcml/mc.py:466-626``), compiles that text unchanged
with gcc behind ``clshim.h`` and runs ``McKernel`` through ``ref_driver.c``.

The rendered C text is a reference source and never enters the repository: it is
written to a temporary directory and only the compiled ``.so`` is kept under
``oracle/_ref/`` (git-ignored, but shipped to the GPU box like any built ``.so``).
"""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')

CFLAGS_EXACT = ['-O2', '-ffp-contract=off']
CFLAGS_FAST = ['-O3', '-march=native', '-ffast-math']

GEOMETRY_ID = {'mcml': 0, 'mcvox': 1, 'mccyl': 2}


class XoRefArgs(ctypes.Structure):
    _fields_ = [
        ('num_packets', ctypes.c_uint32),
        ('num_packets_done', ctypes.c_void_p),
        ('num_kernels', ctypes.c_void_p),
        ('rmax', ctypes.c_double),      # passed on in the kernel's precision (ref_driver.c)
        ('rng_x', ctypes.c_void_p),
        ('rng_a', ctypes.c_void_p),
        ('g0', ctypes.c_uint32),
        ('g1', ctypes.c_void_p),
        ('g2', ctypes.c_void_p),
        ('g3', ctypes.c_void_p),
        ('source', ctypes.c_void_p),
        ('surface', ctypes.c_void_p),
        ('trace', ctypes.c_void_p),
        ('fluence', ctypes.c_void_p),
        ('detectors', ctypes.c_void_p),
        ('fp_lut', ctypes.c_void_p),
        ('int_buffer', ctypes.c_void_p),
        ('float_buffer', ctypes.c_void_p),
        ('accumulator_buffer', ctypes.c_void_p),
        ('trace_float_stride', ctypes.c_uint64),
        ('trace_int_stride', ctypes.c_uint64),
    ]


def compile_rendered(src: str, geometry: str, name: str = None,
                     cflags=None, outdir: str = REF_DIR, stable: bool = False) -> str:
    """gcc-compile rendered reference kernel text; returns the .so path.
    ``stable``: fixed file name ``libref_<name>.so`` (bench baselines)."""
    cflags = list(CFLAGS_EXACT if cflags is None else cflags)
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(HERE, 'clshim.h'), 'rb') as f:
        shim = f.read()
    with open(os.path.join(HERE, 'ref_driver.c'), 'rb') as f:
        drv = f.read()
    digest = hashlib.sha1(
        src.encode() + shim + drv + ' '.join(cflags).encode() +
        geometry.encode()).hexdigest()[:16]
    so = os.path.join(outdir, 'libref_{}_{}.so'.format(name or geometry, digest))
    if stable:
        so = os.path.join(outdir, 'libref_{}.so'.format(name))
    elif os.path.exists(so):
        return so
    with tempfile.TemporaryDirectory() as tmp:
        ksrc = os.path.join(tmp, 'kernel.c')
        with open(ksrc, 'w') as f:
            f.write('#include "clshim.h"\n')
            f.write(src)
        # the SamplingVolume kernel is rendered into every source but compiled
        # only under MC_USE_SAMPLING_VOLUME (set by Trace, mcsv.template.c)
        has_sv = ['-DXO_REF_HAS_SV'] if re.search(
            r'^\s*#define\s+MC_USE_SAMPLING_VOLUME\s+(TRUE|1)\b', src, re.M) else []
        cmd = ['gcc', '-std=gnu11', '-fgnu89-inline', '-w', '-fPIC', '-shared',
               '-pthread', '-I', HERE, '-DXO_REF_GEOMETRY=%d' % GEOMETRY_ID[geometry]
               ] + has_sv + cflags + [ksrc, os.path.join(HERE, 'ref_driver.c'),
                             '-o', so + '.tmp', '-lm']
        subprocess.check_call(cmd)
    os.replace(so + '.tmp', so)
    return so


def _addr(obj):
    if obj is None:
        return None
    if isinstance(obj, np.ndarray):
        return obj.ctypes.data
    return ctypes.addressof(obj)


class RefKernel:
    """The reference ``McKernel`` of one reference ``Mc`` object, on the CPU."""

    def __init__(self, mc_obj, geometry: str, name: str = None, cflags=None,
                 outdir: str = REF_DIR):
        self.mc = mc_obj
        self.geometry = geometry
        self.name = name
        self.cflags = cflags
        self.outdir = outdir
        self._lib = None
        self._src_hash = None
        # precision of the rendered kernel (McDataTypesDouble: binary64 buffers, and the
        # cl_khr_fp64 macro an OpenCL compiler would predefine)
        self.np_float = np.dtype(getattr(mc_obj.types, 'np_float', 'float32'))
        if self.np_float.itemsize == 8:
            self.cflags = list(CFLAGS_EXACT if cflags is None else cflags) + \
                ['-DXO_REF_DOUBLE', '-Dcl_khr_fp64=1']

    # -- reference host side ------------------------------------------------
    def pack(self, nphotons: int):
        m = self.mc
        m._pack(int(nphotons))
        if m._cl_src is None:
            m._build_src()
        if self._lib is None:
            self.so_path = compile_rendered(
                m._cl_src, self.geometry, self.name, self.cflags, self.outdir)
            self._lib = ctypes.CDLL(self.so_path)
            for fn in ('xo_ref_run_static', 'xo_ref_run_dynamic'):
                getattr(self._lib, fn).argtypes = [
                    ctypes.POINTER(XoRefArgs), ctypes.c_uint32]
                getattr(self._lib, fn).restype = None
        return m._packed

    def packed_bytes(self) -> dict:
        """Raw bytes of every packed struct (for host-mirror parity tests)."""
        out = {}
        for key, val in self.mc._packed.items():
            if val is None:
                continue
            if isinstance(val, np.ndarray):
                out[key] = val.tobytes()
            else:
                out[key] = bytes(memoryview(val).cast('B'))
        return out

    def fp_lut(self) -> np.ndarray:
        mgr = self.mc.float_r_lut_manager
        if len(mgr) == 0:
            return np.zeros(1, self.np_float)
        return np.ascontiguousarray(mgr.pack_into(None), dtype=self.np_float)

    # -- run ------------------------------------------------------------------
    def run(self, nphotons: int, nthreads: int, schedule: str = 'static',
            rng_x: np.ndarray = None):
        """Returns dict(accu, ints, floats, rng_x, num_kernels, done)."""
        m = self.mc
        P = self.pack(nphotons)
        nphotons = int(nphotons)
        accu = np.zeros(max(int(m.cl_rw_accumulator_allocator.size), 1), np.uint64)
        ints = np.zeros(max(int(m.cl_rw_int_allocator.size), 1), np.int32)
        floats = np.zeros(max(int(m.cl_rw_float_allocator.size), 1), self.np_float)
        lut = self.fp_lut()
        done = np.zeros(1, np.uint32)
        nk = np.zeros(1, np.uint32)
        x = (m.rng_seeds_x if rng_x is None else rng_x).copy()
        a = np.ascontiguousarray(m.rng_seeds_a)
        dummy = np.zeros(4, np.uint64)

        args = XoRefArgs()
        args.num_packets = nphotons
        args.num_packets_done = done.ctypes.data
        args.num_kernels = nk.ctypes.data
        args.rmax = float(self.np_float.type(m.rmax))
        args.rng_x = x.ctypes.data
        args.rng_a = a.ctypes.data
        keep = []
        if self.geometry == 'mcvox':
            vox = np.ascontiguousarray(m.voxels.data(m))
            keep.append(vox)
            args.g0 = len(m.materials)
            args.g1 = _addr(P['voxels'])
            args.g2 = vox.ctypes.data
            args.g3 = _addr(P['materials'])
        else:
            args.g0 = len(m.layers)
            args.g1 = _addr(P['layers'])
        args.source = _addr(P['source'])
        args.surface = _addr(P.get('surface_layouts')) or dummy.ctypes.data
        args.trace = _addr(P.get('trace')) or dummy.ctypes.data
        args.fluence = _addr(P.get('fluence')) or dummy.ctypes.data
        args.detectors = _addr(P.get('detectors')) or dummy.ctypes.data
        args.fp_lut = lut.ctypes.data
        args.int_buffer = ints.ctypes.data
        args.float_buffer = floats.ctypes.data
        args.accumulator_buffer = accu.ctypes.data
        trace = getattr(m, 'trace', None)
        if trace is not None and schedule == 'static':
            args.trace_float_stride = 8*int(trace.maxlen)
            args.trace_int_stride = 1
        fn = {'static': self._lib.xo_ref_run_static,
              'dynamic': self._lib.xo_ref_run_dynamic}[schedule]
        fn(ctypes.byref(args), int(nthreads))
        return dict(accu=accu, ints=ints, floats=floats, rng_x=x,
                    num_kernels=int(nk[0]), done=int(done[0]), lut=lut)

    def sampling_volume(self, sv, trace_n: np.ndarray, trace_data: np.ndarray):
        """The reference's ``SamplingVolume`` kernel on trace rows (``trace_n``:
        int32[n], ``trace_data``: float32[n*maxlen*8]).  Host part as in
        mcml/mc.py:1040-1215: the allocators are cleared, the trace and the
        sampling volume are packed again and the rows are uploaded."""
        m = self.mc
        trace = m.trace
        n = int(np.asarray(trace_n).size)
        m.cl_rw_accumulator_allocator.clear()
        m.cl_rw_float_allocator.clear()
        m.cl_rw_int_allocator.clear()
        tp = trace.cl_pack(m, None, nphotons=n)
        sp = sv.cl_pack(m, None)
        accu = np.zeros(max(int(m.cl_rw_accumulator_allocator.size), 1), np.uint64)
        ibuf = np.zeros(max(int(m.cl_rw_int_allocator.size), 1), np.int32)
        fbuf = np.zeros(max(int(m.cl_rw_float_allocator.size), 1), self.np_float)
        _, do, co, _ = np.frombuffer(bytes(memoryview(tp).cast('B')), np.uint32)[:4].tolist()
        ibuf[co:co + n] = trace_n
        flat = np.asarray(trace_data, self.np_float).reshape(-1)
        fbuf[do:do + flat.size] = flat
        total = np.zeros(1, np.uint64)
        fn = self._lib.xo_ref_run_sv
        fn.argtypes = [ctypes.c_uint32] + [ctypes.c_void_p]*6
        fn.restype = None
        fn(n, ctypes.addressof(tp), ctypes.addressof(sp), total.ctypes.data,
           ibuf.ctypes.data, fbuf.ctypes.data, accu.ctypes.data)
        return dict(accu=accu, total_weight=int(total[0]),
                    packed_sv=bytes(memoryview(sp).cast('B')),
                    packed_sv_trace=bytes(memoryview(tp).cast('B')))
