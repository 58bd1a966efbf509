"""Device runtime shared by the simulators (counterpart of
``xopto/mcbase/mcworker.py``: ClWorker + buffer/LUT/RNG mixins).

Holds the CUDA context/stream (through ``libxopto_b200.so``), the three flat
read-write buffers with their allocators, the read-only float LUT pool, the
per-thread MWC seeds, and the NVRTC kernel cache.  Everything on the device goes
through :mod:`pyxopto_b200.cu.abi`; there is no other execution path.
"""
import hashlib
import os
import time

import numpy as np

from .. import KERNEL_PATH, KERNEL_CACHE_PATH, VERBOSE
from ..cl import clinfo, clrng
from ..cu import abi
from . import mctypes
from .mcutil.buffer import BufferAllocator
from .mcutil.lut import LutManager

# the safe-prime table holds 500 000 multipliers; one seeds the seed generator
MAX_SEEDS = 499999
TARGET_ARCH = 'sm_100a'             # B200; the only architecture this engine targets

_header_cache = None


def kernel_headers() -> dict:
    """All ``csrc/kernels/*.cuh`` files as in-memory NVRTC headers."""
    global _header_cache
    if _header_cache is None:
        hdrs = {}
        for name in sorted(os.listdir(KERNEL_PATH)):
            if name.endswith('.cuh'):
                with open(os.path.join(KERNEL_PATH, name)) as f:
                    hdrs[name] = f.read()
        _header_cache = hdrs
    return _header_cache


def nvrtc_options(deterministic: bool, extra=()) -> list:
    opts = ['--std=c++17', '-lineinfo', '-default-device']
    if deterministic:
        opts += ['--fmad=false', '--prec-div=true', '--prec-sqrt=true', '--ftz=false']
    else:
        opts += ['--use_fast_math', '--extra-device-vectorization']
    return opts + list(extra)


def compile_kernel(src: str, deterministic: bool, arch: str = TARGET_ARCH,
                   extra_options=(), use_cache: bool = True):
    """TU text -> (cubin bytes, log, cache_hit).  Works without a GPU."""
    headers = kernel_headers()
    options = nvrtc_options(deterministic, extra_options)
    h = hashlib.sha256()
    h.update(src.encode())
    h.update(arch.encode())
    h.update('\0'.join(options).encode())
    for name in sorted(headers):
        h.update(name.encode())
        h.update(headers[name].encode())
    key = h.hexdigest()[:32]
    path = os.path.join(KERNEL_CACHE_PATH, '{}_{}.cubin'.format(arch, key))
    if use_cache and os.path.exists(path):
        with open(path, 'rb') as f:
            return f.read(), '', True
    cubin, log = abi.compile_cubin(src, 'xo_kernel.cu', arch, options, headers)
    if use_cache:
        try:
            os.makedirs(KERNEL_CACHE_PATH, exist_ok=True)
            tmp = path + '.tmp{}'.format(os.getpid())
            with open(tmp, 'wb') as f:
                f.write(cubin)
            os.replace(tmp, path)
        except OSError:
            pass
    return cubin, log, False


class _LaneBuffers(dict):
    """Named device buffers of a simulator.  A sweep runs consecutive configurations on two
    *lanes* (mcsweep.Sweep): everything a launch writes or re-uploads - accumulators,
    counters, MWC states, packed medium / plugin tables - exists once per lane, so the kernel
    of configuration k + 1 may start while configuration k drains; the big read-only inputs
    (voxel maps) are shared.  Code addresses buffers by their plain names; the lane of the
    owner picks the instance."""
    SHARED = ('voxel_data', 'voxel_packed')

    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def _key(self, name):
        lane = self._owner._lane
        if lane and isinstance(name, str) and name not in self.SHARED:
            return '{}#{}'.format(name, lane)
        return name

    def __getitem__(self, name):
        return super().__getitem__(self._key(name))

    def __setitem__(self, name, value):
        super().__setitem__(self._key(name), value)

    def __contains__(self, name):
        return super().__contains__(self._key(name))

    def get(self, name, default=None):
        return super().get(self._key(name), default)

    def pop(self, name, *default):
        return super().pop(self._key(name), *default)


class CuWorker:
    """Base class of the ``Mc`` simulators."""
    _lane = 0                        # see _LaneBuffers

    def __init__(self, types=mctypes.McDataTypesSingle, cl_devices=None,
                 cl_build_options=None, cl_profiling: bool = False, rnginit=None):
        self._types = types
        self._cl_device_arg = cl_devices
        self._cl_build_options = list(cl_build_options or [])
        self._cl_profiling = bool(cl_profiling)
        self._ctx = None
        self._lane_streams = {}          # lane -> (stream, (event, event))
        self._lane_seeds = {}            # lane -> (seeds_x, seeds_a) of the lanes above 0
        self._cl_buffers = _LaneBuffers(self)
        self._np_buffers = {}
        self._allocators = {
            'accumulator': BufferAllocator(types.np_accu),
            'float': BufferAllocator(types.np_float),
            'int': BufferAllocator(types.np_int),
        }
        self._float_lut = LutManager(types.np_float)
        self._rng = clrng.Random()
        self._rng_seeds_x, self._rng_seeds_a = self._rng.seeds(MAX_SEEDS, xinit=rnginit)
        self._rnginit = rnginit
        self._seeds_on_device = set()    # lanes whose MWC states are on the device
        self._modules = {}
        self._pinned_downloads = {}

    # -- device ---------------------------------------------------------------
    def _device_ordinal(self) -> int:
        d = self._cl_device_arg
        if d is None:
            return int(os.environ.get('LOCAL_RANK', 0)) if os.environ.get(
                'XOPTO_DEVICE_FROM_RANK') else 0
        if isinstance(d, clinfo.Device):
            return d.ordinal
        if isinstance(d, (int, np.integer)):
            return int(d)
        if isinstance(d, str):
            return clinfo.device(d).ordinal
        if isinstance(d, (list, tuple)) and d:
            first = d[0]
            return first.ordinal if isinstance(first, clinfo.Device) else int(first)
        raise TypeError('Unsupported cl_devices argument {!r}'.format(d))

    def _ensure_device(self):
        if self._ctx is None:
            self._ctx = abi.Context(self._device_ordinal())
        return self._ctx

    def _lane_state(self):
        st = self._lane_streams.get(self._lane)
        if st is None:
            self._ensure_device()
            st = (abi.Stream(self._ctx), (abi.Event(self._ctx), abi.Event(self._ctx)))
            self._lane_streams[self._lane] = st
        return st

    # the stream / timing events of the current lane
    _stream = property(lambda self: self._lane_state()[0])
    _events = property(lambda self: self._lane_state()[1])

    cl_context = property(lambda self: self._ensure_device())
    cl_queue = property(lambda self: (self._ensure_device(), self._stream)[1])
    cl_device = property(lambda self: self._ensure_device().info)
    types = property(lambda self: self._types)
    rng = property(lambda self: self._rng)
    rng_seeds_x = property(lambda self: self._rng_seeds_x)
    rng_seeds_a = property(lambda self: self._rng_seeds_a)
    cl_buffers = property(lambda self: self._cl_buffers)
    np_buffers = property(lambda self: self._np_buffers)

    @property
    def cl_max_threads(self) -> int:
        return (MAX_SEEDS//512)*512          # 499 712, mcworker.py:1227

    # -- allocators / LUTs ------------------------------------------------------
    cl_rw_accumulator_allocator = property(lambda self: self._allocators['accumulator'])
    cl_rw_float_allocator = property(lambda self: self._allocators['float'])
    cl_rw_int_allocator = property(lambda self: self._allocators['int'])
    float_r_lut_manager = property(lambda self: self._float_lut)

    def cl_allocate_rw_accumulator_buffer(self, owner, shape, download=True):
        return self._allocators['accumulator'].allocate(owner, shape, download)

    def cl_allocate_rw_float_buffer(self, owner, shape, download=True):
        return self._allocators['float'].allocate(owner, shape, download)

    def cl_allocate_rw_int_buffer(self, owner, shape, download=True):
        return self._allocators['int'].allocate(owner, shape, download)

    def append_r_lut(self, data: np.ndarray, force: bool = False):
        return self._float_lut.append(data, force)

    def clear_r_luts(self):
        self._float_lut.clear()

    def _clear_allocations(self):
        for a in self._allocators.values():
            a.clear()
        self.clear_r_luts()

    # -- buffers ----------------------------------------------------------------
    def _buffer(self, name: str, nbytes: int) -> abi.Buffer:
        """Named device buffer of at least ``nbytes`` (re-allocated on growth)."""
        self._ensure_device()
        buf = self._cl_buffers.get(name)
        if buf is None or buf.size < nbytes:
            if buf is not None:
                buf.release()
            buf = abi.Buffer(self._ctx, max(int(nbytes), 16))
            self._cl_buffers[name] = buf
        return buf

    def _buffer_with_headroom(self, name: str, nbytes: int) -> abi.Buffer:
        """Like :meth:`_buffer`, but a (re)allocation reserves 1.5 x ``nbytes``."""
        buf = self._cl_buffers.get(name)
        if buf is not None and buf.size >= nbytes:
            return buf
        return self._buffer(name, nbytes + nbytes//2)

    def cl_r_buffer(self, name: str, host=None, size: int = None) -> abi.Buffer:
        """Upload a packed struct / array into a named read-only buffer."""
        if host is None:
            return self._buffer(name, size or 16)
        if isinstance(host, np.ndarray):
            host = np.ascontiguousarray(host)
            nbytes = host.nbytes
        else:
            nbytes = len(bytes(memoryview(host).cast('B')))
        buf = self._buffer(name, nbytes)
        # pageable host memory is staged by the driver before the call returns,
        # so the asynchronous form is safe for the (small) packed structs
        buf.upload(self._stream, host, blocking=not self._async_uploads)
        return buf

    # sweeps double-buffer the accumulators on the device (mcsweep.Sweep)
    _accumulator_slot = 0
    _async_uploads = False

    def _rw_name(self, kind: str) -> str:
        if kind == 'accumulator' and self._accumulator_slot:
            return 'rw_accumulator_{}'.format(self._accumulator_slot)
        return 'rw_' + kind

    def _counters_name(self) -> str:
        return 'counters_{}'.format(self._accumulator_slot) if self._accumulator_slot \
            else 'counters'

    def _rw_flat_buffer(self, kind: str, fill: bool = True) -> abi.Buffer:
        alloc = self._allocators[kind]
        nbytes = max(alloc.size, 1)*alloc.dtype.itemsize
        buf = self._buffer(self._rw_name(kind), nbytes)
        if fill:
            buf.fill(self._stream, 0, np.uint32, count=(nbytes + 3)//4)
        return buf

    def _upload_seeds(self, copy: bool = False):
        lane = self._lane
        if lane not in self._seeds_on_device or copy:
            if lane == 0:
                x, a = self._rng_seeds_x, self._rng_seeds_a
            else:
                # a lane above 0 draws from its own seed set (hashed initializer, as the
                # ranks of a sharded run do): two lanes never share a stream
                if lane not in self._lane_seeds:
                    from .. import parallel
                    base = self._rnginit if self._rnginit is not None else \
                        int(self._rng_seeds_x[0])
                    self._lane_seeds[lane] = self._rng.seeds(
                        MAX_SEEDS, xinit=parallel.seed_for_rank(base, lane << 20))
                x, a = self._lane_seeds[lane]
            self.cl_r_buffer('rng_seeds_x', x)
            self.cl_r_buffer('rng_seeds_a', a)
            self._seeds_on_device.add(lane)

    PINNED_MIN_BYTES = 1 << 20

    def _download_host_array(self, kind: str, a) -> np.ndarray:
        """Host array for the download of allocation ``a``.  Large accumulator
        grids (201^3 x 8 B = 65 MB) come back through a page-locked staging array
        that is reused by every run: ``update_data()`` only reads the raw integers
        (raw += accu*(1/k)), it never keeps the array."""
        if kind != 'accumulator' or a.size*np.dtype(a.dtype).itemsize < self.PINNED_MIN_BYTES:
            return np.empty(a.shape, dtype=a.dtype)
        key = (kind, a.offset, a.shape)
        host = self._pinned_downloads.get(key)
        if host is None:
            if len(self._pinned_downloads) >= 4:
                self._pinned_downloads.clear()
            host = abi.pinned_empty(self._ctx, a.shape, a.dtype)
            self._pinned_downloads[key] = host
        return host

    def _download_allocations(self, owner, nphotons: int):
        """{dtype: [ndarray per allocation]} for one plugin (cf. mc.py:1020-1038)."""
        out = {}
        for kind, alloc in self._allocators.items():
            buf = self._cl_buffers.get('rw_' + kind)
            for a in alloc.allocations(owner):
                if not a.download:
                    continue
                if hasattr(owner, 'np_buffer'):
                    host = owner.np_buffer(self, a, nphotons=nphotons)
                else:
                    host = self._download_host_array(kind, a)
                buf.download(self._stream, host, offset=a.offset*alloc.dtype.itemsize)
                out.setdefault(alloc.dtype, []).append(host)
        return out

    # -- kernels ----------------------------------------------------------------
    def _module(self, src: str, deterministic: bool):
        """Build (or fetch from cache) and load the module for a TU."""
        self._ensure_device()
        key = hashlib.sha1(src.encode()).hexdigest() + str(deterministic)
        mod = self._modules.get(key)
        if mod is None:
            t0 = time.perf_counter()
            cubin, log, hit = compile_kernel(
                src, deterministic, arch=self._ctx.arch,
                extra_options=self._cl_build_options)
            mod = abi.Module(self._ctx, cubin=cubin)
            mod.build_log = log
            mod.build_time = time.perf_counter() - t0
            mod.cache_hit = hit
            if VERBOSE:
                print('pyxopto_b200: kernel {} in {:.3f} s'.format(
                    'loaded from cache' if hit else 'built', mod.build_time))
            self._modules[key] = mod
        return mod

    def launch_geometry(self, kernel, block: int, dynamic_shared: int,
                        maxthreads: int = None):
        """(grid, block): one resident wave of CTAs over all SMs, capped by the
        number of available MWC seeds (499 712) and ``maxthreads``."""
        info = self._ctx.info
        per_sm = max(kernel.occupancy(block, dynamic_shared), 1)
        grid = info['multiprocessor_count']*per_sm
        limit = self.cl_max_threads if maxthreads is None else \
            min(int(maxthreads), self.cl_max_threads)
        grid = max(1, min(grid, limit//block))
        return grid, block
