"""ctypes binding of libxopto_b200.so (include/xopto_b200.h).

This is the layer that stands where ``pyopencl`` stands in the reference
(``xopto/mcbase/mcworker.py``): contexts, queues (streams), programs (NVRTC
modules), buffers, copies, fills, launches and events, all on NumPy host
buffers.  Failures raise :class:`RuntimeError` carrying the library message -
the same convention as ``pyopencl.RuntimeError`` in the reference.

There is no fallback: if the shared library is missing the import fails, and if
no CUDA driver/GPU is present every device call raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, 'libxopto_b200.so')
SRC_PATH = os.path.join(_PKG_DIR, 'csrc', 'xo_abi.cpp')
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), 'include', 'xopto_b200.h')

XO_ERR_NO_DRIVER = -1
XO_ERR_COMPILE = -3

Handle = ctypes.c_uint64


class DeviceInfo(ctypes.Structure):
    _fields_ = [
        ('name', ctypes.c_char*256),
        ('cc_major', ctypes.c_int32), ('cc_minor', ctypes.c_int32),
        ('multiprocessor_count', ctypes.c_int32),
        ('max_threads_per_block', ctypes.c_int32),
        ('max_threads_per_multiprocessor', ctypes.c_int32),
        ('max_shared_per_block_optin', ctypes.c_int32),
        ('regs_per_multiprocessor', ctypes.c_int32),
        ('clock_rate_khz', ctypes.c_int32),
        ('l2_cache_bytes', ctypes.c_int32),
        ('reserved', ctypes.c_int32),
        ('total_global_mem', ctypes.c_uint64),
    ]


class Arg(ctypes.Structure):
    _fields_ = [
        ('kind', ctypes.c_int32), ('size', ctypes.c_int32),
        ('value', ctypes.c_void_p), ('buffer', Handle), ('offset', ctypes.c_uint64),
    ]


def build_library(force: bool = False) -> str:
    """Compile csrc/xo_abi.cpp in-tree (g++; no CUDA link-time dependency)."""
    if not force and os.path.exists(LIB_PATH) and \
            os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(SRC_PATH),
                                              os.path.getmtime(HEADER_PATH)):
        return LIB_PATH
    cuda_home = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    cmd = ['g++', '-std=c++17', '-O2', '-Wall', '-fPIC', '-shared',
           '-I', os.path.join(cuda_home, 'include'), SRC_PATH,
           '-o', LIB_PATH + '.tmp', '-ldl', '-pthread']
    subprocess.check_call(cmd)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


_SIGNATURES = {
    'xo_last_error': (ctypes.c_char_p, []),
    'xo_version': (ctypes.c_int, []),
    'xo_device_count': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32)]),
    'xo_device_get_info': (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(DeviceInfo)]),
    'xo_ctx_create': (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(Handle)]),
    'xo_ctx_destroy': (ctypes.c_int, [Handle]),
    'xo_stream_create': (ctypes.c_int, [Handle, ctypes.POINTER(Handle)]),
    'xo_stream_destroy': (ctypes.c_int, [Handle]),
    'xo_stream_sync': (ctypes.c_int, [Handle]),
    'xo_stream_native': (ctypes.c_int, [Handle, ctypes.POINTER(ctypes.c_uint64)]),
    'xo_module_build': (ctypes.c_int, [
        Handle, ctypes.c_char_p, ctypes.c_char_p,
        ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
        ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
        ctypes.POINTER(Handle), ctypes.c_char_p, ctypes.c_size_t]),
    'xo_compile': (ctypes.c_int, [
        ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
        ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
        ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
        ctypes.POINTER(Handle), ctypes.c_char_p, ctypes.c_size_t]),
    'xo_blob_size': (ctypes.c_int, [Handle, ctypes.POINTER(ctypes.c_size_t)]),
    'xo_blob_copy': (ctypes.c_int, [Handle, ctypes.c_void_p, ctypes.c_size_t]),
    'xo_blob_free': (ctypes.c_int, [Handle]),
    'xo_module_load': (ctypes.c_int, [Handle, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.POINTER(Handle)]),
    'xo_module_unload': (ctypes.c_int, [Handle]),
    'xo_module_get_kernel': (ctypes.c_int, [Handle, ctypes.c_char_p,
                                            ctypes.POINTER(Handle)]),
    'xo_kernel_get_attributes': (ctypes.c_int, [Handle] + [ctypes.POINTER(ctypes.c_int32)]*4),
    'xo_kernel_occupancy': (ctypes.c_int, [Handle, ctypes.c_int32, ctypes.c_size_t,
                                           ctypes.POINTER(ctypes.c_int32)]),
    'xo_buffer_alloc': (ctypes.c_int, [Handle, ctypes.c_size_t, ctypes.POINTER(Handle)]),
    'xo_buffer_free': (ctypes.c_int, [Handle]),
    'xo_buffer_size': (ctypes.c_int, [Handle, ctypes.POINTER(ctypes.c_size_t)]),
    'xo_buffer_device_ptr': (ctypes.c_int, [Handle, ctypes.POINTER(ctypes.c_uint64)]),
    'xo_copy_h2d': (ctypes.c_int, [Handle, Handle, ctypes.c_size_t, ctypes.c_void_p,
                                   ctypes.c_size_t, ctypes.c_int32]),
    'xo_copy_d2h': (ctypes.c_int, [Handle, ctypes.c_void_p, Handle, ctypes.c_size_t,
                                   ctypes.c_size_t, ctypes.c_int32]),
    'xo_fill': (ctypes.c_int, [Handle, Handle, ctypes.c_size_t, ctypes.c_size_t,
                               ctypes.c_int32, ctypes.c_void_p]),
    'xo_host_alloc': (ctypes.c_int, [Handle, ctypes.c_size_t,
                                     ctypes.POINTER(ctypes.c_void_p)]),
    'xo_host_free': (ctypes.c_int, [Handle, ctypes.c_void_p]),
    'xo_launch': (ctypes.c_int, [Handle, Handle, ctypes.c_uint32, ctypes.c_uint32,
                                 ctypes.c_uint32, ctypes.POINTER(Arg), ctypes.c_int32]),
    'xo_event_create': (ctypes.c_int, [Handle, ctypes.POINTER(Handle)]),
    'xo_event_destroy': (ctypes.c_int, [Handle]),
    'xo_event_record': (ctypes.c_int, [Handle, Handle]),
    'xo_event_sync': (ctypes.c_int, [Handle]),
    'xo_event_elapsed_ms': (ctypes.c_int, [Handle, Handle, ctypes.POINTER(ctypes.c_float)]),
    'init_RNG': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_uint32, ctypes.c_uint64]),
}

_lib = None


def lib():
    """The loaded shared library (built on first use when sources are newer)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build_library()
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


class CudaError(RuntimeError):
    """Raised for every failed library call (cf. pyopencl.RuntimeError)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def check(code: int):
    if code != 0:
        msg = lib().xo_last_error().decode('utf-8', 'replace')
        raise CudaError(code, 'libxopto_b200: ' + msg)


def device_count() -> int:
    n = ctypes.c_int32(0)
    check(lib().xo_device_count(ctypes.byref(n)))
    return n.value


def device_info(ordinal: int = 0) -> dict:
    info = DeviceInfo()
    check(lib().xo_device_get_info(ordinal, ctypes.byref(info)))
    out = {name: getattr(info, name) for name, _ in DeviceInfo._fields_}
    out['name'] = info.name.decode()
    return out


def _cstr_array(items):
    arr = (ctypes.c_char_p*max(len(items), 1))()
    for i, it in enumerate(items):
        arr[i] = it.encode() if isinstance(it, str) else it
    return arr


def compile_cubin(src: str, name: str, arch: str, options, headers: dict):
    """NVRTC compile without a device; returns (cubin bytes, log)."""
    blob = Handle(0)
    log = ctypes.create_string_buffer(1 << 16)
    hn = list(headers.keys())
    hs = [headers[k] for k in hn]
    check(lib().xo_compile(
        src.encode(), name.encode(), arch.encode(),
        _cstr_array(options), len(options),
        _cstr_array(hs), _cstr_array(hn), len(hn),
        ctypes.byref(blob), log, len(log)))
    size = ctypes.c_size_t(0)
    check(lib().xo_blob_size(blob, ctypes.byref(size)))
    data = ctypes.create_string_buffer(size.value)
    check(lib().xo_blob_copy(blob, data, size.value))
    check(lib().xo_blob_free(blob))
    return data.raw, log.value.decode('utf-8', 'replace')


class Context:
    """cl.Context equivalent: the primary CUDA context of one device."""

    def __init__(self, ordinal: int = 0):
        self.ordinal = int(ordinal)
        h = Handle(0)
        check(lib().xo_ctx_create(self.ordinal, ctypes.byref(h)))
        self.handle = h
        self.info = device_info(self.ordinal)

    @property
    def arch(self) -> str:
        cc = '%d%d' % (self.info['cc_major'], self.info['cc_minor'])
        return 'sm_' + cc + ('a' if self.info['cc_major'] >= 9 else '')

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                lib().xo_ctx_destroy(self.handle)
        except Exception:
            pass


class Stream:
    """cl.CommandQueue equivalent (in-order CUDA stream)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = Handle(0)
        check(lib().xo_stream_create(ctx.handle, ctypes.byref(h)))
        self.handle = h

    def synchronize(self):
        check(lib().xo_stream_sync(self.handle))

    @property
    def native(self) -> int:
        v = ctypes.c_uint64(0)
        check(lib().xo_stream_native(self.handle, ctypes.byref(v)))
        return v.value

    def __del__(self):
        try:
            lib().xo_stream_destroy(self.handle)
        except Exception:
            pass


class Event:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = Handle(0)
        check(lib().xo_event_create(ctx.handle, ctypes.byref(h)))
        self.handle = h

    def record(self, stream: Stream):
        check(lib().xo_event_record(self.handle, stream.handle))
        return self

    def synchronize(self):
        check(lib().xo_event_sync(self.handle))

    def elapsed_ms(self, later: 'Event') -> float:
        ms = ctypes.c_float(0)
        check(lib().xo_event_elapsed_ms(self.handle, later.handle, ctypes.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib().xo_event_destroy(self.handle)
        except Exception:
            pass


class Buffer:
    """cl.Buffer equivalent: a device allocation of ``size`` bytes."""

    def __init__(self, ctx: Context, size: int):
        self.ctx = ctx
        self.size = int(size)
        h = Handle(0)
        check(lib().xo_buffer_alloc(ctx.handle, self.size, ctypes.byref(h)))
        self.handle = h

    @property
    def device_ptr(self) -> int:
        v = ctypes.c_uint64(0)
        check(lib().xo_buffer_device_ptr(self.handle, ctypes.byref(v)))
        return v.value

    def upload(self, stream: Stream, host, offset: int = 0, blocking: bool = True):
        if isinstance(host, (bytes, bytearray)):
            # (a read-only array view keeps the bytes alive for the whole call; an
            # asynchronous upload of a temporary object is the caller's business)
            host = np.frombuffer(host, dtype=np.uint8)
        raw, nbytes = _host_ptr(host)
        check(lib().xo_copy_h2d(stream.handle, self.handle, offset, raw, nbytes,
                                int(blocking)))

    def download(self, stream: Stream, host: np.ndarray, offset: int = 0,
                 blocking: bool = True):
        raw, nbytes = _host_ptr(host)
        check(lib().xo_copy_d2h(stream.handle, raw, self.handle, offset, nbytes,
                                int(blocking)))
        return host

    def fill(self, stream: Stream, value, dtype, count: int = None, offset: int = 0):
        pattern = np.array([value], dtype=dtype)
        if count is None:
            count = (self.size - offset)//pattern.itemsize
        check(lib().xo_fill(stream.handle, self.handle, offset, int(count),
                            pattern.itemsize, pattern.ctypes.data))

    def release(self):
        if self.handle:
            lib().xo_buffer_free(self.handle)
            self.handle = Handle(0)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class _PinnedBlock:
    """Owner of one page-locked host allocation (xo_host_alloc)."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p(0)
        check(lib().xo_host_alloc(ctx.handle, max(self.nbytes, 1), ctypes.byref(p)))
        self.ptr = p.value
        self.ctypes_array = (ctypes.c_ubyte*max(self.nbytes, 1)).from_address(self.ptr)

    def __del__(self):
        try:
            if self.ptr:
                lib().xo_host_free(self.ctx.handle, ctypes.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


def pinned_empty(ctx: Context, shape, dtype) -> np.ndarray:
    """NumPy array in page-locked host memory: device <-> host copies of it run
    at full PCIe rate and can be asynchronous.  The allocation lives as long as
    the array (or any view of it) does."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    block = _PinnedBlock(ctx, n*dtype.itemsize)
    flat = np.frombuffer(block.ctypes_array, dtype=np.uint8, count=n*dtype.itemsize)
    arr = flat.view(dtype).reshape(shape)
    # np.frombuffer keeps `block.ctypes_array` alive; tie the block to it
    block.ctypes_array._xo_block = block
    return arr


def _host_ptr(host):
    if isinstance(host, np.ndarray):
        if not host.flags['C_CONTIGUOUS']:
            raise ValueError('host array must be C-contiguous')
        return host.ctypes.data, host.nbytes
    if isinstance(host, (bytes, bytearray)):
        raise TypeError('pass bytes through numpy.frombuffer(): a temporary copy would be '
                        'freed before the copy engine reads it')
    return ctypes.addressof(host), ctypes.sizeof(host)


class Kernel:
    def __init__(self, module: 'Module', name: str):
        self.module = module
        self.name = name
        h = Handle(0)
        check(lib().xo_module_get_kernel(module.handle, name.encode(), ctypes.byref(h)))
        self.handle = h

    def attributes(self) -> dict:
        vals = [ctypes.c_int32(0) for _ in range(4)]
        check(lib().xo_kernel_get_attributes(self.handle, *[ctypes.byref(v) for v in vals]))
        return dict(zip(('num_regs', 'static_shared', 'max_threads', 'local_bytes'),
                        (v.value for v in vals)))

    def occupancy(self, block: int, dynamic_shared: int = 0) -> int:
        n = ctypes.c_int32(0)
        check(lib().xo_kernel_occupancy(self.handle, block, dynamic_shared, ctypes.byref(n)))
        return n.value

    def launch(self, stream: Stream, grid: int, block: int, args, dynamic_shared: int = 0):
        """``args``: Buffer | (Buffer, byte_offset) | NumPy scalar | ctypes
        struct | bytes (passed by value)."""
        n = len(args)
        arr = (Arg*max(n, 1))()
        keep = []
        for i, a in enumerate(args):
            if isinstance(a, Buffer):
                arr[i].kind, arr[i].buffer, arr[i].offset = 1, a.handle, 0
            elif isinstance(a, tuple) and isinstance(a[0], Buffer):
                arr[i].kind, arr[i].buffer, arr[i].offset = 1, a[0].handle, int(a[1])
            else:
                if isinstance(a, np.generic):
                    a = np.array([a])
                if isinstance(a, np.ndarray):
                    a = np.ascontiguousarray(a)
                    keep.append(a)
                    arr[i].value, arr[i].size = a.ctypes.data, a.nbytes
                elif isinstance(a, (bytes, bytearray)):
                    b = ctypes.create_string_buffer(bytes(a), len(a))
                    keep.append(b)
                    arr[i].value, arr[i].size = ctypes.addressof(b), len(a)
                else:
                    keep.append(a)
                    arr[i].value, arr[i].size = ctypes.addressof(a), ctypes.sizeof(a)
                arr[i].kind = 0
        check(lib().xo_launch(stream.handle, self.handle, int(grid), int(block),
                              int(dynamic_shared), arr, n))


class Module:
    """cl.Program(...).build() equivalent: an NVRTC-compiled, loaded module."""

    def __init__(self, ctx: Context, src: str = None, name: str = 'xo_kernel.cu',
                 options=(), headers: dict = None, cubin: bytes = None):
        self.ctx = ctx
        self.log = ''
        h = Handle(0)
        if cubin is not None:
            self._image = ctypes.create_string_buffer(cubin, len(cubin))
            check(lib().xo_module_load(ctx.handle, self._image, len(cubin), ctypes.byref(h)))
        else:
            headers = headers or {}
            hn = list(headers.keys())
            hs = [headers[k] for k in hn]
            log = ctypes.create_string_buffer(1 << 16)
            options = list(options)
            code = lib().xo_module_build(
                ctx.handle, src.encode(), name.encode(),
                _cstr_array(options), len(options),
                _cstr_array(hs), _cstr_array(hn), len(hn),
                ctypes.byref(h), log, len(log))
            self.log = log.value.decode('utf-8', 'replace')
            check(code)
        self.handle = h
        self._kernels = {}

    def kernel(self, name: str) -> Kernel:
        if name not in self._kernels:
            self._kernels[name] = Kernel(self, name)
        return self._kernels[name]

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return self.kernel(name)

    def __del__(self):
        try:
            lib().xo_module_unload(self.handle)
        except Exception:
            pass


def init_rng(fora: np.ndarray, n_rng: int, xinit: int):
    """Binding of ``init_RNG`` (same contract as rng.cpp:64-103)."""
    fora = np.ascontiguousarray(fora, dtype=np.uint32)
    if fora.size < n_rng + 1:
        raise ValueError('need n_rng + 1 multipliers')
    x = np.zeros(n_rng, dtype=np.uint64)
    a = np.zeros(n_rng, dtype=np.uint32)
    rc = lib().init_RNG(x.ctypes.data, a.ctypes.data, fora.ctypes.data, n_rng,
                        ctypes.c_uint64(int(xinit)))
    if rc != 0:
        raise ValueError('Invalid xinit value for the given set of multipliers!')
    return x, a
