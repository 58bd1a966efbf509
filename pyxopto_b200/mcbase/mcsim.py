"""Simulator base class: kernel assembly + ``run()`` (counterpart of the shared
parts of ``xopto/mc{ml,vox,cyl}/mc.py``: ``_pack``, ``_build_src``, ``run``).

Kernel assembly: where the reference renders one OpenCL-C translation unit from
plugin text fragments (mcbase/mcsrc.py + jinja2, mc.py:531-626), this engine
emits a ~30-line CUDA translation unit that (1) maps the resolved compile-time
options onto ``XO_*`` defines, (2) binds the plugin slots to the hand-written
CUDA structs of ``csrc/kernels`` with typedefs, (3) pins the struct layouts with
``static_assert(sizeof(...))`` against the ctypes structs (the compile-time twin
of the reference's ``sizeof_datatypes`` kernel, mc.py:674-728) and (4) includes
the geometry kernel header.  NVRTC compiles it for sm_100a; cubins are cached
in-tree by content hash.
"""
import ctypes
import os
import time
import weakref

import numpy as np

from . import mcoptions, mctypes
from . import mcsv as mcsv_module
from .mcworker import CuWorker, compile_kernel

DEFAULT_BLOCK = 256
# with a fluence grid one 1024-thread CTA per SM shares a single large window of
# the grid in shared memory (xo::FluWindow)
FLUENCE_BLOCK = 1024
SHARED_RESERVE = 2048            # static shared memory + driver reserve per CTA
PRIVATE_BINS_MAX = 4096          # detector bins privatised per CTA (32 KB)
LUT_SHARED_MAX_BYTES = 64*1024   # pf lookup tables staged in shared memory


def _fragment(obj, attr, mc) -> str:
    """OpenCL-C text of a plugin attribute that is a string or a callable(mc)."""
    frag = getattr(obj, attr, None)
    if frag is None:
        return ''
    if callable(frag):
        frag = frag(mc)
    return str(frag or '')


def _c_float(v: float, double: bool = False) -> str:
    """Literal in the kernels' floating-point type (FP_LITERAL of the reference)."""
    if double:
        v = float(v)
        if np.isinf(v):
            return 'XO_INF' if v > 0 else '(-XO_INF)'
        return repr(v)
    v = float(np.float32(v))
    if np.isinf(v):
        return '__int_as_float(0x7f800000)' if v > 0 else '__int_as_float(0xff800000)'
    return '{!r}f'.format(v)


class _ResidentRows:
    """Stand-in for a Trace whose rows only exist on the device."""

    def __init__(self, dev: dict, trace):
        self.nphotons, self.maxlen = dev['n'], dev['maxlen']
        if dev.get('token') is None:
            dev['token'] = object()
        self._device_token = dev['token']
        self._trace = trace

    def cl_pack(self, mc, target=None, nphotons=None):
        return self._trace.cl_pack(mc, target, nphotons=nphotons)


class McBase(CuWorker):
    """Plugin bookkeeping, option resolution, TU emission and the run loop."""

    kernel_header = None     # e.g. 'mcml_kernel.cuh'
    supports_surface_layouts = False
    geometry = None          # 'mcml' | 'mcvox' | 'mccyl'

    def __init__(self, source, detectors=None, trace=None, fluence=None,
                 surface=None, types=mctypes.McDataTypesSingle, options=None,
                 rnginit=None, cl_devices=None, cl_build_options=None,
                 cl_profiling: bool = False):
        super().__init__(types=types, cl_devices=cl_devices,
                         cl_build_options=cl_build_options,
                         cl_profiling=cl_profiling, rnginit=rnginit)
        if surface is not None and not self.supports_surface_layouts:
            raise NotImplementedError(
                'Surface layouts are only part of the accelerated path of the '
                'layered simulator (mcml).')
        # plugin objects built with the reference package are rebuilt with this
        # package's classes from their own description (pyxopto_b200/adopt.py)
        from ..adopt import adopt
        source, detectors, trace, fluence, surface = (
            adopt(obj, self.geometry) for obj in (source, detectors, trace, fluence, surface))
        self._surface = surface
        self._source = source
        self._detectors = detectors
        self._trace = trace
        self._fluence = fluence
        self._options = list(options or [])
        self._rmax = float('inf')
        self._packed = {}
        self._run_report = {}
        self._reduce_hook = None
        self._last_src = None
        # plugin types are frozen at construction (mc.py:344-389, 486-529)
        self._obj_types = {
            'source': type(source),
            'detectors': None if detectors is None else detectors.types(),
            'fluence': type(fluence), 'trace': type(trace),
            'surface': None if surface is None else surface.types()}

    # -- user-facing properties -------------------------------------------------
    def _set_rmax(self, r):
        self._rmax = float(r)

    rmax = property(lambda self: self._rmax, _set_rmax, None,
                    'Packets farther than rmax from the source are terminated.')
    source = property(lambda self: self._source)
    detectors = property(lambda self: self._detectors)
    trace = property(lambda self: self._trace)
    fluence = property(lambda self: self._fluence)
    surface = property(lambda self: self._surface)
    run_report = property(lambda self: self._run_report)
    options = property(lambda self: self._options)

    # -- options ----------------------------------------------------------------
    def _plugin_option_lists(self):
        lists = [self._types.cl_options(self), self._options]
        for obj in (self._source, self._surface, self._detectors, self._fluence, self._trace):
            if obj is not None:
                lists.append(obj.fetch_cl_options(self))
        return lists

    def resolved_options(self) -> dict:
        opts = mcoptions.resolve_cl_options(*self._plugin_option_lists())
        if opts.get('MC_USE_DOUBLE_PRECISION'):
            # binary64 runs use the reference-structured loops only (static schedule,
            # reference expression order, no contraction): csrc/kernels/xo_math_double.cuh
            opts['XO_DETERMINISTIC'] = True
        return opts

    @property
    def deterministic(self) -> bool:
        return bool(self.resolved_options().get('XO_DETERMINISTIC', False))

    # -- packing ----------------------------------------------------------------
    def _check_types(self):
        if type(self._source) is not self._obj_types['source']:
            raise ValueError('The photon packet source type/kind must not '
                             'change between simulation calls!')
        if self._detectors is not None and \
                self._detectors.types() != self._obj_types['detectors']:
            raise ValueError('Detector types must not change between simulation calls!')

    def _pack_medium(self):
        raise NotImplementedError

    def _pack(self, nphotons: int):
        """Assign buffer offsets (pack order: medium, source, detectors, fluence,
        trace - mc.py:466-529) and fill the packed structs."""
        self._clear_allocations()
        self._check_types()
        self._pack_medium()
        self._packed['source'], _, _ = self._source.cl_pack(
            self, self._packed.get('source'))
        if self._surface is not None:
            if self._surface.types() != self._obj_types['surface']:
                raise ValueError('Surface layout types must not change between '
                                 'simulation calls!')
            self._packed['surface_layouts'] = self._surface.cl_pack(self, self._packed.get('surface_layouts'))
        if self._detectors is not None:
            self._packed['detectors'] = self._detectors.cl_pack(
                self, self._packed.get('detectors'))
        if self._fluence is not None:
            self._packed['fluence'] = self._fluence.cl_pack(
                self, self._packed.get('fluence'))
        if self._trace is not None:
            self._packed['trace'] = self._trace.cl_pack(
                self, self._packed.get('trace'), int(nphotons))
        return self._packed

    # -- translation unit -------------------------------------------------------
    def _plugin_bindings(self):
        """[(typedef name, CUDA struct, packed ctypes struct or None)]"""
        raise NotImplementedError

    def _detector_bindings(self):
        from ..mcml import mcdetector
        dets = self._detectors
        out = []
        for loc, Name in (('top', 'XoDetTop'), ('bottom', 'XoDetBottom'),
                          ('specular', 'XoDetSpecular')):
            det = getattr(dets, loc) if dets is not None else mcdetector.DetectorDefault()
            out.append((Name, det.fetch_cu_type(self), det.fetch_cl_type(self)))
        return out

    def kernel_source(self, block: int = DEFAULT_BLOCK, min_blocks: int = 1) -> str:
        """The CUDA translation unit for the current plugin set / options."""
        opts = self.resolved_options()
        trace_flags = int(opts.get('MC_USE_TRACE', 0))
        is_double = bool(opts.get('MC_USE_DOUBLE_PRECISION', False))
        lines = [
            '// generated by pyxopto_b200 ({})'.format(self.geometry),
            '#define XO_DOUBLE {}'.format(int(bool(opts.get('MC_USE_DOUBLE_PRECISION', False)))),
            '#define XO_DETERMINISTIC {}'.format(int(bool(opts.get('XO_DETERMINISTIC', False)))),
            '#define XO_METHOD {}'.format(int(opts.get('MC_METHOD', 0))),
            '#define XO_USE_LOTTERY {}'.format(int(bool(opts.get('MC_USE_LOTTERY', True)))),
            '#define XO_ENHANCED_RNG {}'.format(int(bool(opts.get('MC_USE_ENHANCED_RNG', False)))),
            '#define XO_WEIGHT_MIN {}'.format(
                _c_float(opts.get('MC_PACKET_WEIGHT_MIN', 1e-4), is_double)),
            '#define XO_LOTTERY_CHANCE {}'.format(
                _c_float(opts.get('MC_PACKET_LOTTERY_CHANCE', 0.1), is_double)),
            '#define XO_TRACE {}'.format(trace_flags),
            '#define XO_USE_EVENTS {}'.format(int(bool(opts.get('MC_USE_EVENTS', False)))),
            '#define XO_TRACK_OPL {}'.format(
                int(bool(opts.get('MC_TRACK_OPTICAL_PATHLENGTH', False)))),
            '#define XO_FLUENCE_RATE {}'.format(
                int(bool(opts.get('MC_FLUENCE_MODE_RATE', False)))),
            '#define XO_TRACE_ALIGNED {}'.format(int(self._trace_aligned())),
            '#define XO_TRACE_STORE_HINT {}'.format(int(self.trace_store_hint)),
            '#define XO_USE_RMAX {}'.format(int(self._rmax_needed())),
            '#define XO_FLU_WINDOW {}'.format(int(self._window_enabled())),
            '#define XO_PF_G0 {}'.format(int(self._pf_isotropic_possible())),
            '#define XO_DEPOSIT_MAGIC {}'.format(int(self._deposit_magic(opts))),
            '#define XO_BLOCK {}'.format(int(block)),
            '#define XO_MIN_BLOCKS {}'.format(int(min_blocks)),
        ]
        lines += self._extra_defines(opts)
        bindings = self._plugin_bindings()
        user = self._user_fragments(bindings)
        if user:
            # compile-time options as the macros OpenCL-C fragments test (#if MC_USE_...)
            for key in sorted(opts):
                if (key.startswith('MC_') or key == 'TRACE_ENTRY_LEN') and \
                        isinstance(opts[key], (bool, int, np.integer)):
                    lines.append('#define {} {}'.format(key, int(opts[key])))
            lines += ['#define {} 1'.format(self._USER_ADAPTERS[name][1]) for name, _ in user]
        lines += ['#include "xo_core.cuh"', '#include "xo_pf.cuh"',
                  '#include "xo_detectors.cuh"', '#include "xo_fluence.cuh"']
        lines += self._extra_includes()
        if user:
            lines.append('#include "xo_clcompat.cuh"')
            for name, obj in user:
                lines.append('// ---- declarations of the user plugin bound to {} ({})'.format(
                    name, type(obj).__name__))
                lines.append(_fragment(obj, 'cl_declaration', self))
            lines.append('#include "xo_clcompat_slots.cuh"')
        checks = []
        user_names = {name for name, _ in user}
        for name, cu_type, cl_type in bindings:
            if name in user_names:
                cu_type = self._USER_ADAPTERS[name][0]
            lines.append('typedef {} {};'.format(cu_type, name))
            if cl_type is not None:
                size = ctypes.sizeof(cl_type)
                if opts.get('MC_USE_DOUBLE_PRECISION') and size % 8:
                    # a host struct declared with _pack_ = 1 (mccyl FiZ) has no tail
                    # padding; the device struct is naturally aligned - the member
                    # offsets agree and the enclosing struct is checked as a whole
                    size += 8 - size % 8
                checks.append('static_assert(sizeof({}) == {}, "{} layout differs from '
                              'the packed host struct");'.format(name, size, name))
        if user:
            if self.clcompat_geometry_header:
                lines.append('#include "{}"'.format(self.clcompat_geometry_header))
            for name, obj in user:
                lines.append('// ---- implementation of the user plugin bound to {}'.format(name))
                if name == 'XoSource':
                    # a launch only *requests* the specular deposit (xo_clcompat.cuh)
                    lines.append('#define mcsim_specular_detector_deposit XO_CLC_SPECULAR_REQUEST')
                lines.append(_fragment(obj, 'cl_implementation', self))
                if name == 'XoSource':
                    lines.append('#undef mcsim_specular_detector_deposit')
            lines.append('#include "xo_clcompat_glue.cuh"')
        lines.append('#include "{}"'.format(self.kernel_header))
        lines += checks
        lines += self._extra_checks()
        return '\n'.join(lines) + '\n'

    def _extra_defines(self, opts):
        return []

    # -- user-written plugins (OpenCL-C fragments, csrc/kernels/xo_clcompat*.cuh) --
    # slot typedef -> (adapter struct, flag macro)
    _USER_ADAPTERS = {
        'XoPf': ('xo::PfUser', 'XO_USER_PF'),
        'XoSource': ('xo::SrcUser', 'XO_USER_SOURCE'),
        'XoDetTop': ('xo::DetUserTop', 'XO_USER_DET_TOP'),
        'XoDetBottom': ('xo::DetUserBottom', 'XO_USER_DET_BOTTOM'),
        'XoDetSpecular': ('xo::DetUserSpecular', 'XO_USER_DET_SPECULAR'),
        'XoDetOuter': ('xo::DetUserOuter', 'XO_USER_DET_OUTER'),
        'XoFluence': ('xo::FluUser', 'XO_USER_FLUENCE'),
        'XoTrace': ('xo::TraceUser', 'XO_USER_TRACE'),
        'XoSurfTop': ('xo::SurfUserTop', 'XO_USER_SURF_TOP'),
        'XoSurfBottom': ('xo::SurfUserBottom', 'XO_USER_SURF_BOTTOM'),
    }
    user_plugin_slots = ('XoPf',)        # slots of this geometry that take fragments
    clcompat_geometry_header = None

    def _plugin_objects(self) -> dict:
        """{slot typedef name: plugin object} for the slots that may hold a
        user-written plugin."""
        return {}

    def _user_fragments(self, bindings):
        """[(slot, object)] of the plugins that have no hand-written CUDA struct
        (``cu_type`` is None) but carry the reference's OpenCL-C fragment protocol
        (``cl_declaration`` / ``cl_implementation``, mcobject.py:29-170)."""
        objs = self._plugin_objects()
        out = []
        for name, cu_type, _ in bindings:
            if cu_type is not None:
                continue
            obj = objs.get(name)
            if obj is None or not hasattr(obj, 'cl_implementation') or \
                    name not in self.user_plugin_slots:
                raise NotImplementedError(
                    'The plugin bound to {} ({}) has neither a CUDA implementation '
                    '(cu_type) nor OpenCL-C fragments this geometry can compile.'.format(
                        name, type(obj).__name__ if obj is not None else None))
            out.append((name, obj))
        if self._user_trace():
            # (the kernel headers bind XoTrace themselves: no typedef from the bindings)
            if 'XoTrace' not in self.user_plugin_slots:
                raise NotImplementedError(
                    'This simulator has no slot for a user-written trace '
                    '({}).'.format(type(self._trace).__name__))
            out.append(('XoTrace', self._trace))
        return out

    def _user_trace(self) -> bool:
        """True when the trace object carries its own OpenCL-C fragments
        (``cl_implementation``, mctrace.py:541-585) instead of the built-in event record."""
        return self._trace is not None and hasattr(self._trace, 'cl_implementation')

    def _scattering_pfs(self):
        """Phase functions of the layers / materials a packet can scatter in."""
        return None

    def _pf_isotropic_possible(self) -> bool:
        """False when every scattering phase function has a packed anisotropy
        ``g != 0``: the kernel then drops the isotropic special case of Hg / MHg
        (hg.py:74-90 draws a third number when g == 0)."""
        pfs = self._scattering_pfs()
        if not pfs:
            return True
        for pf in pfs:
            g = getattr(pf, 'g', None)
            if g is None:
                return True
            try:
                if float(np.float32(g)) == 0.0:
                    return True
            except (TypeError, ValueError):
                return True
        return False

    def _deposit_magic(self, opts) -> bool:
        """True when every fluence deposit of the throughput loop is below 2^23 - 1
        in fixed point (weight <= 1, deposition mode, k <= 0x7FFFFF): the kernel then
        converts with one FFMA + LOP3 instead of the conversion unit."""
        flu = self._fluence
        if flu is None or opts.get('MC_FLUENCE_MODE_RATE', False) or \
                getattr(flu, 'cu_type', None) is None:
            return False
        wmin = float(opts.get('MC_PACKET_WEIGHT_MIN', 1e-4))
        chance = float(opts.get('MC_PACKET_LOTTERY_CHANCE', 0.1))
        return int(getattr(flu, 'k', 1 << 30)) <= 0x7FFFFF and wmin <= chance and \
            getattr(self._source, 'cu_type', None) is not None

    def _rmax_needed(self) -> bool:
        """False when the rmax test can never fire (compiled out of the loop)."""
        return bool(np.isfinite(np.float32(self._rmax)))

    # waiting lanes per warp (interface pending / packet needed) that trigger a
    # service round of the throughput loops; None: the source's / geometry's default
    refill_lanes = None
    default_refill_lanes = 1
    min_blocks = None                # __launch_bounds__ second argument (None: automatic)

    def _min_blocks(self, block: int) -> int:
        if self.min_blocks is not None:
            return int(self.min_blocks)
        # a full trace is a store stream (one 32-byte event per trip and lane): more
        # resident warps keep more stores in flight (measured C4, 1e6 packets: 2.69 ms
        # at 3 CTAs of 256 / 70 registers, 2.85 ms at 4 / 64, 2.95 ms at 2 / 76)
        if self.geometry == 'mcml' and self._trace is not None and block <= 256 and \
                not self.deterministic:
            return 3
        return 1
    chunk_max = 16

    def _refill_lanes(self) -> int:
        if self.refill_lanes is not None:
            return int(min(max(self.refill_lanes, 1), 32))
        return int(getattr(self._source, 'cu_refill_lanes', self.default_refill_lanes))

    def _loop_name(self) -> str:
        """Which photon loop of the kernel header the launch ran (run_report['loop'])."""
        return 'reference-structured' if self.deterministic else 'throughput'

    def _queue_bytes(self, block: int) -> int:
        """Shared memory of the throughput loops behind the window: per-warp launch
        queues (32 slots of 40 bytes)."""
        return 40*block + 16

    def _extra_includes(self):
        return []

    def _extra_checks(self):
        return []

    # cache hints of the 256-bit event store (developer knob; 0: plain store)
    trace_store_hint = int(os.environ.get('XOPTO_TRACE_STORE_HINT', '0'))

    def _trace_aligned(self) -> int:
        """Alignment class of the trace rows in the float buffer: 2 = 32 bytes (an
        event leaves with one 256-bit store), 1 = 16 bytes (two 128-bit stores),
        0 = scalar stores."""
        tr = self._packed.get('trace')
        if self._trace is None or tr is None or self._user_trace():
            return 0
        off = int(tr.data_buffer_offset)
        if np.dtype(self._types.np_float).itemsize != 4:
            return 0                    # binary64 events: plain stores
        wide = os.environ.get('XOPTO_TRACE_STORE', '256') == '256'     # developer knob
        return 2 if (off % 8 == 0 and wide) else (1 if off % 4 == 0 else 0)

    def export_src(self, filename: str = None, nphotons: int = 1) -> str:
        self._pack(nphotons)
        src = self.kernel_source()
        if filename:
            with open(filename, 'w') as f:
                f.write(src)
        return src

    def compile(self, nphotons: int = 1, block: int = DEFAULT_BLOCK, min_blocks: int = 1,
                arch: str = 'sm_100a'):
        """NVRTC-compile the kernel for the current configuration without a
        device (used by ``__graft_entry__.build`` and the CPU test-suite)."""
        self._pack(nphotons)
        src = self.kernel_source(block, min_blocks)
        self._last_src = src
        return compile_kernel(src, self.deterministic, arch=arch,
                              extra_options=self._cl_build_options)

    # -- launch ----------------------------------------------------------------
    def _source_focus(self):
        pos = getattr(self._source, 'position', None)
        if pos is None:
            return (0.0, 0.0, 0.0)
        return tuple(float(v) for v in np.asarray(pos, dtype=np.float64).reshape(-1)[:3])

    fluence_window_bytes = None      # None: automatic
    # packets per kernel launch: the device packet counter is 32 bits wide and every
    # thread claims at most one more chunk (<= chunk_max packets, mcml / mcvox
    # throughput loops: one per lane) after the budget is exhausted, so the counter
    # ends at most max_threads*chunk_max (< 2^24) above the budget and must not wrap
    max_batch = 0xFFFFFFFF - (1 << 24)
    _keep_accumulators = False       # batches of one run: do not re-zero the accumulators
    _packet_counter_start = 0
    fluence_block = FLUENCE_BLOCK

    def _window_enabled(self) -> bool:
        return self._fluence is not None and hasattr(self._fluence, 'cu_window') and \
            self.fluence_window_bytes != 0

    def _fluence_window(self, block: int, base_bytes: int):
        """xo::FluWindow (6 x uint32) for the current fluence plugin."""
        none = np.zeros(6, dtype=np.uint32)
        if not self._window_enabled():
            return none
        budget = self.fluence_window_bytes
        if budget is None:
            cap = self._ctx.info['max_shared_per_block_optin'] if self._ctx is not None \
                else 227*1024
            if block >= 512:
                # 1024 threads per SM: one CTA of 1024 or two of 512
                budget = cap//(1024//min(block, 1024)) - SHARED_RESERVE - base_bytes
            else:
                budget = 40*1024 - base_bytes
        bins = int(max(budget, 0))//4
        win = self._fluence.cu_window(self, bins, self._source_focus())
        win = np.asarray(win, dtype=np.uint32)
        if int(win[3])*int(win[4])*int(win[5]) > bins:
            return none
        return win

    def _shared_layout(self, medium_bytes: int):
        """(dynamic shared bytes, lut floats staged, private bins)."""
        lut_len = 0
        if len(self._float_lut) and self._float_lut.size*4 <= LUT_SHARED_MAX_BYTES and \
                np.dtype(self._types.np_float).itemsize == 4:
            # (binary64 tables are read from global memory: the shared-memory layout of
            # the kernels counts 4-byte words)
            lut_len = self._float_lut.size
        priv_len = 0
        if self._detectors is not None:
            det_bins = 0
            for det in self._detectors:
                for a in self.cl_rw_accumulator_allocator.allocations(det):
                    det_bins = max(det_bins, a.offset + a.size)
            priv_len = min(det_bins, PRIVATE_BINS_MAX)
        words = (medium_bytes//4 + 3) & ~3
        if lut_len:
            words += (lut_len + 3) & ~3
        words += 2*priv_len
        words = ((words + 3) & ~3) + 4      # 16 aligned bytes in front of the fluence window
        return words*4 + 16, lut_len, priv_len

    def _trace_tails_unread(self) -> bool:
        """True when nothing ever reads trace events beyond a packet's own count:
        the float buffer holds only trace rows and they leave the device through
        the device-side filter, whose compaction writes the zero tails itself
        (TraceCompact), or feed ``sampling_volume``, which reads ``n`` events per
        packet.  The reference zero-fills the whole trace buffer before every run
        (16 GB for 1e6 packets x 512 events); here that pass is skipped then."""
        tr = self._trace
        if tr is None or tr.filter is None or not self._device_filter_applies():
            return False
        if self.resolved_options().get('MC_USE_EVENTS', False):
            # with an event mask a packet can record nothing at all: the filter then
            # reads slot maxlen - 1 of its row (numpy's index -1), which must be zero
            return False
        allocs = self._allocators['float'].allocations()
        return bool(allocs) and all(a.owner is tr for a in allocs)

    def _kernel_args(self, nphotons, bufs, lut_len, priv_len, chunk, refill, window):
        raise NotImplementedError

    def _medium_bytes(self) -> int:
        raise NotImplementedError

    def _upload_medium(self):
        raise NotImplementedError

    def _packed_or_dummy(self, key, size=8):
        p = self._packed.get(key)
        if p is None:
            return bytes(size)
        return p

    def run(self, nphotons: int, out=None, wgsize: int = None, maxthreads: int = None,
            copyseeds: bool = False, exportsrc: str = None, verbose: bool = False,
            download: bool = True, synchronize: bool = True):
        """Simulate ``nphotons`` packets.  Returns ``(trace, fluence, detectors)``
        result objects like the reference (mc.py:730-1018); with ``out`` the new
        data are accumulated into the given previous results."""
        nphotons = int(nphotons)
        if nphotons > self._types.mc_cnt_max:
            raise ValueError('Maximum number of photon packets that can be '
                             'simulated in a single run is limited to {:,d}!'.format(
                                 self._types.mc_cnt_max))
        if nphotons > self.max_batch:
            # 64-bit packet counter family (McDataTypesSingleCnt64): the device
            # counter stays 32 bits wide, the host runs the budget in batches that
            # continue the MWC streams.  The batches accumulate where they are: the
            # 64-bit integer accumulators stay on the device from batch to batch
            # (exact sums) and come to the host once, after the last one
            if self._trace is not None:
                raise ValueError('A traced run is limited to {:,d} packets!'.format(self.max_batch))
            if not (download and synchronize):
                raise ValueError('Runs of more than {:,d} packets need download=True, '
                                 'synchronize=True'.format(self.max_batch))
            remaining = nphotons
            kernel_ms = iterations = 0
            hook = self._reduce_hook
            try:
                while remaining > 0:
                    batch = min(remaining, self.max_batch)
                    # (multi-GPU: the ranks combine their totals once, after the last batch)
                    self._reduce_hook = hook if batch == remaining else None
                    self.run(batch, wgsize=wgsize, maxthreads=maxthreads, copyseeds=copyseeds,
                             exportsrc=exportsrc, verbose=verbose, download=False)
                    kernel_ms += self._run_report['kernel_ms']
                    iterations += self._run_report['iterations']
                    remaining -= batch
                    self._keep_accumulators = True      # the next batch adds on top
            finally:
                self._keep_accumulators = False
                self._reduce_hook = hook
            results = self._collect_results(nphotons, out)
            self._run_report.update(items=nphotons, kernel_ms=kernel_ms, iterations=iterations)
            return results
        t0 = time.perf_counter()
        self._ensure_device()
        self._device_trace = None        # rows of the previous run are overwritten
        if not synchronize and (download or copyseeds):
            raise ValueError('synchronize=False requires download=False, copyseeds=False')
        self._async_uploads = not synchronize
        self._pack(nphotons)
        deterministic = self.deterministic
        if wgsize:
            block = int(wgsize)
        elif self._fluence is not None and not deterministic:
            block = self.fluence_block
        else:
            block = DEFAULT_BLOCK
        src = self.kernel_source(block=block, min_blocks=self._min_blocks(block))
        self._last_src = src
        if exportsrc:
            with open(exportsrc, 'w') as f:
                f.write(src)
        mod = self._module(src, deterministic)
        kernel = mod.kernel('McKernel')
        t_build = time.perf_counter()

        # uploads (mc.py:840-884)
        self._upload_seeds(copy=False)
        counters = np.zeros(4, dtype=np.uint32)   # done, kernels, iterations (u64)
        # (test hook: a throughput-mode launch that starts with the packet counter
        # at S simulates the packets [S, nphotons) - exercises the top of the
        # 32-bit counter range without simulating 4e9 packets)
        counters[0] = self._packet_counter_start
        cbuf = self.cl_r_buffer(self._counters_name(), counters)
        self._upload_medium()
        if len(self._float_lut):
            lut_host = self._float_lut.pack_into(None).astype(self._types.np_float)
        else:
            lut_host = np.zeros(4, self._types.np_float)
        lbuf = self.cl_r_buffer('fp_lut', lut_host)
        abuf = self._rw_flat_buffer('accumulator', fill=not self._keep_accumulators)
        fbuf = self._rw_flat_buffer('float', fill=not self._trace_tails_unread())
        ibuf = self._rw_flat_buffer('int')
        shared, lut_len, priv_len = self._shared_layout(self._medium_bytes())
        queue_bytes = 0 if deterministic else self._queue_bytes(block)
        window = self._fluence_window(block, shared + queue_bytes)
        shared += 4*int(window[3])*int(window[4])*int(window[5]) + queue_bytes
        grid, block = self.launch_geometry(kernel, block, shared, maxthreads)
        nthreads = grid*block
        if deterministic:
            chunk = 0
        else:
            # packets claimed per atomic: small enough that the last chunk of a
            # lane (the load-imbalance tail) is < 2 % of its work
            chunk = int(min(max(nphotons//(nthreads*64), 1), self.chunk_max))
        bufs = dict(counters=cbuf, lut=lbuf, accu=abuf, floats=fbuf, ints=ibuf,
                    rng_x=self._cl_buffers['rng_seeds_x'],
                    rng_a=self._cl_buffers['rng_seeds_a'])
        refill = 1 if deterministic else self._refill_lanes()
        args = self._kernel_args(nphotons, bufs, lut_len, priv_len, chunk, refill, window)
        t_up = time.perf_counter()

        ev0, ev1 = self._events
        ev0.record(self._stream)
        kernel.launch(self._stream, grid, block, args, dynamic_shared=shared)
        ev1.record(self._stream)
        if self._reduce_hook is not None:
            self._reduce_hook(self, abuf, self.cl_rw_accumulator_allocator.size)
        self._async_uploads = False
        if not synchronize:
            # queued only (mcsweep.Sweep): the caller orders its own copies on
            # the stream and waits on its own events
            self._run_report = {'items': nphotons, 'grid': grid, 'block': block,
                                'launched_threads': nthreads, 'cache_hit': mod.cache_hit}
            return (None, None, None)
        self._stream.synchronize()
        t_exec = time.perf_counter()
        kernel_ms = ev0.elapsed_ms(ev1)

        cbuf.download(self._stream, counters)
        if copyseeds:
            self._cl_buffers['rng_seeds_x'].download(self._stream, self._rng_seeds_x)

        results = (None, None, None)
        if download:
            results = self._collect_results(nphotons, out)
        t_down = time.perf_counter()
        self._run_report = {
            'upload': t_up - t_build, 'build': t_build - t0,
            'execution': t_exec - t_up, 'download': t_down - t_exec,
            'kernel_ms': kernel_ms, 'threads': int(counters[1]),
            'iterations': int(counters[2:4].view(np.uint64)[0]),
            'launched_threads': nthreads, 'grid': grid, 'block': block,
            'shared_bytes': shared, 'private_bins': priv_len, 'lut_shared': lut_len,
            'chunk': chunk, 'refill': refill, 'fluence_window': [int(v) for v in window], 'items': nphotons, 'cache_hit': mod.cache_hit,
            'kernel_attributes': kernel.attributes(), 'loop': self._loop_name(),
        }
        if verbose:
            print('pyxopto_b200 run: build {:.3f} s, upload {:.3f} s, kernel {:.3f} ms, '
                  'download {:.3f} s, {} threads'.format(
                      t_build - t0, t_up - t_build, kernel_ms, t_down - t_exec, nthreads))
        return results

    def _collect_results(self, nphotons, out):
        out_trace = out_fluence = out_detectors = None
        if out is not None:
            out_trace, out_fluence, out_detectors = out
        trace_res = fluence_res = detectors_res = None
        if self._trace is not None:
            trace_res = out_trace if out_trace is not None else type(self._trace)(self._trace)
            if self._trace.filter is not None and self._device_filter_applies():
                # (binary64 rows go through the host filter, like the reference's)
                lazy = out_trace is None and self.lazy_trace_rows
                n_sel, rows, n_dropped = self.filter_trace_on_device(
                    nphotons, download=not lazy)
                data = {np.dtype(self._types.np_float): [rows],
                        np.dtype(self._types.np_int): [n_sel]}
                trace_res.update_data(self, data, nphotons=nphotons, prefiltered=True,
                                      n_dropped=n_dropped)
                if lazy:
                    # the accepted rows (16 KB each at maxlen 512) stay on the device
                    # until somebody reads `trace.data`
                    trace_res._set_lazy_rows(self._lazy_rows_loader(
                        (n_sel.size, int(self._trace.maxlen)), rows.dtype))
                    self._lazy_trace = weakref.ref(trace_res)
            else:
                data = self._download_allocations(self._trace, nphotons)
                trace_res.update_data(self, data, nphotons=nphotons)
                tp = self._packed['trace']
                self._device_trace = dict(
                    n=nphotons, maxlen=int(self._trace.maxlen),
                    ints=(self._cl_buffers['rw_int'], 4*int(tp.count_buffer_offset)),
                    floats=(self._cl_buffers['rw_float'], 4*int(tp.data_buffer_offset)))
            if out_trace is None and (self._trace.filter is None or self._device_filter_applies()):
                # the rows of this result are still on the device
                token = object()
                self._device_trace['token'] = token
                trace_res._device_token = token
        if self._fluence is not None:
            fluence_res = out_fluence if out_fluence is not None \
                else type(self._fluence)(self._fluence)
            if self.lazy_fluence and self._fluence_resident_add(fluence_res, nphotons):
                pass
            else:
                self._collect_fluence(fluence_res, nphotons)
        if self._detectors is not None:
            detectors_res = out_detectors if out_detectors is not None \
                else type(self._detectors)(self._detectors)
            for det, res in zip(self._detectors, detectors_res):
                data = self._download_allocations(det, nphotons)
                if data:
                    detectors_res.update_data(self, res, data, nphotons=nphotons)
        return trace_res, fluence_res, detectors_res

    def _collect_fluence(self, fluence_res, nphotons):
        """The reference's flow: the accumulators of this run reach the host and are
        added to the result there (fluence.py:349-376)."""
        scaled = self._download_scaled_fluence(fluence_res)
        if scaled is not None:
            owned = False
            if isinstance(scaled, tuple):
                scaled, owned = scaled
            fluence_res.update_scaled(scaled, nphotons, owned=owned)
        else:
            data = self._download_allocations(self._fluence, nphotons)
            fluence_res.update_data(self, data, nphotons=nphotons)

    # -- device-resident fluence result -----------------------------------
    # True: the float64 grid of a fluence result stays on the device across
    # ``run(out=...)`` calls - every run adds ``accumulators*(1/k)`` to it there
    # (AccuScaleAdd: the same two IEEE operations per cell as the host's
    # ``raw += accumulators*(1/k)``, so the sums are bit-identical) - and comes to the
    # host when the result's ``raw`` / ``data`` is read; False: the reference's flow, one
    # conversion and one download of the whole grid per run.
    lazy_fluence = False
    _flu_resident = None             # (weakref to the result, grid cells)

    def _fluence_resident_add(self, result, nphotons: int) -> bool:
        from . import mcfluence, rngkernel
        flu = self._fluence
        if type(result).update_data is not mcfluence._FluenceBase.update_data or \
                type(result)._data is not mcfluence._FluenceBase._data:
            return False
        allocs = self.cl_rw_accumulator_allocator.allocations(flu)
        if len(allocs) != 1 or not allocs[0].download:
            return False
        a = allocs[0]
        cells = int(a.size)
        cur = self._flu_resident
        owner = cur[0]() if cur is not None else None
        if owner is not result or cur[1] != cells or result._pending is None:
            if owner is not None:
                owner._materialize()
            self._flu_resident = None
        grid_buf = self._buffer('flu_resident', cells*8)
        add = True
        if self._flu_resident is None:
            have = result._store
            if have is not None:
                # a result that already holds host data: it continues on the device
                if have.size != cells:
                    return False
                grid_buf.upload(self._stream, np.ascontiguousarray(have, dtype=np.float64))
                result._store = None
            else:
                add = False
        src = self._cl_buffers[self._rw_name('accumulator')]
        mod = rngkernel._aux_module(self, True)
        mod.kernel('AccuScaleAdd' if add else 'AccuScale').launch(
            self._stream, 4*self._ctx.info['multiprocessor_count'], 512,
            [(src, a.offset*8), grid_buf, np.uint64(cells), np.float64(1.0/result.k)])
        result._nphotons = result._nphotons + nphotons if add else nphotons
        self._flu_resident = (weakref.ref(result), cells)
        result._pending = self._fluence_loader(cells)
        return True

    def _fluence_loader(self, cells: int):
        def load(result):
            cur = self._flu_resident
            if cur is None or cur[0]() is not result:
                return
            host, owned = self._pinned_result(cells)
            self._cl_buffers['flu_resident'].download(self._stream, host)
            result._store = (host if owned else np.array(host)).reshape(result.shape)
            self._flu_resident = None
        return load

    # grids of at least this many cells are converted to float64 on the device
    # (AccuScale): the host then only copies the page-locked result
    SCALE_ON_DEVICE_MIN = 1 << 20

    def _download_scaled_fluence(self, result):
        """float64 ``accumulators*(1/k)`` of the fluence grid computed on the device
        (bit-identical to ``update_data``'s NumPy expression), or None when the
        plugin does not use the stock conversion / the grid is small."""
        from . import mcfluence
        flu = self._fluence
        if type(result).update_data is not mcfluence._FluenceBase.update_data or \
                not hasattr(result, 'update_scaled'):
            return None
        allocs = self.cl_rw_accumulator_allocator.allocations(flu)
        if len(allocs) != 1 or allocs[0].size < self.SCALE_ON_DEVICE_MIN or \
                not allocs[0].download:
            return None
        src = self._cl_buffers[self._rw_name('accumulator')]
        # a fresh result takes the converted grid as its own array
        return self._scale_on_device(src, allocs[0], 1.0/result.k,
                                     want_owned=result.raw is None)

    # Page-locked float64 result buffers (grid conversions done on the device).  A
    # result object may OWN such a buffer (no host copy, no first-touch page faults on
    # 65 MB): the pool hands out views and reuses a buffer once nothing but the pool
    # refers to it any more; at most RESULT_POOL buffers are kept per size.
    RESULT_POOL = 3

    def _pinned_result(self, size: int):
        """(array, owned): a page-locked float64[size]; ``owned`` True when the caller
        may keep it (a pool buffer that is free), False when it is the shared staging
        array that the next conversion overwrites."""
        import sys
        from ..cu import abi
        pool = self._pinned_downloads.setdefault(('result_pool', size), [])
        for buf in pool:
            # (views of a view share the root array as their base: count its referrers -
            # the pool's own array, the local name and the argument)
            root = buf.base if buf.base is not None else buf
            if sys.getrefcount(root) <= 3:
                return buf.view(), True
        if len(pool) < self.RESULT_POOL:
            buf = abi.pinned_empty(self._ctx, (size,), np.float64)
            pool.append(buf)
            return buf.view(), True
        key = ('accu_scaled', size)
        host = self._pinned_downloads.get(key)
        if host is None:
            host = abi.pinned_empty(self._ctx, (size,), np.float64)
            self._pinned_downloads[key] = host
        return host, False

    def _scale_on_device(self, src, a, inv_k: float, want_owned: bool = False):
        """Page-locked float64 array ``accu[a]*inv_k`` computed by AccuScale (with
        ``want_owned``: ``(array, owned)``, see ``_pinned_result``)."""
        from . import rngkernel
        mod = rngkernel._aux_module(self, True)
        out = self._buffer('accu_scaled', a.size*8)
        grid = 4*self._ctx.info['multiprocessor_count']
        mod.kernel('AccuScale').launch(
            self._stream, grid, 512,
            [(src, a.offset*8), out, np.uint64(a.size), np.float64(inv_k)])
        if want_owned:
            host, owned = self._pinned_result(int(a.size))
        else:
            key = ('accu_scaled', a.size)
            host = self._pinned_downloads.get(key)
            if host is None:
                from ..cu import abi
                host = abi.pinned_empty(self._ctx, (a.size,), np.float64)
                self._pinned_downloads[key] = host
            owned = False
        out.download(self._stream, host)
        return (host, owned) if want_owned else host

    # -- device-side trace filter ------------------------------------
    # True: a Trace with a Filter is filtered and compacted on the device and
    # only the accepted rows are downloaded; False: the reference's flow
    # (download every row, filter on the host).  Both give identical results.
    device_trace_filter = True

    def _device_filter_applies(self) -> bool:
        """The device-side filter reads the built-in single-precision event record."""
        return bool(self.device_trace_filter and not self._user_trace() and
                    np.dtype(self._types.np_float).itemsize == 4)
    _FILTER_SRC = '#include "xo_trace_kernels.cuh"\n'

    # True: the rows accepted by the device-side filter are downloaded when the
    # result's `data` is first read (or before the device buffer is reused)
    lazy_trace_rows = True
    _lazy_trace = None

    def _lazy_rows_loader(self, shape, dtype):
        buf, stream = self._device_trace['floats'], self._stream

        def load():
            rows = np.empty(shape, dtype=dtype)
            if rows.size:
                buf.download(stream, rows)
            return rows
        return load

    def _materialize_lazy_trace(self):
        """Download the rows of an outstanding lazy Trace result (called before
        the compact-row buffer is overwritten)."""
        ref, self._lazy_trace = self._lazy_trace, None
        res = ref() if ref is not None else None
        if res is not None and res.rows_on_device_only:
            res.data                                   # noqa: B018  (triggers the download)

    def filter_trace_on_device(self, nphotons: int, download: bool = True):
        """Evaluate ``self.trace.filter`` on the trace rows of the last run where
        they are, compact the accepted rows in packet order and return
        ``(n, rows, n_dropped)`` (``rows``: structured array [accepted, maxlen]).
        With ``download=False`` only the counts are read back and the compact
        rows stay on the device for ``sampling_volume``."""
        from ..cu import abi
        self._materialize_lazy_trace()
        trace = self._trace
        tp = self._packed['trace']
        nphotons = int(nphotons)
        mod = self._module(self._FILTER_SRC, True)
        counts, ranges = trace.filter.cu_pack(trace.plon)
        fcfg = (ctypes.c_uint32*8)(*[int(v) for v in counts])
        rbuf = self.cl_r_buffer('filter_ranges', ranges)
        block = 256
        grid = max((nphotons + block - 1)//block, 1)
        flags = self._buffer('filter_flags', 4*max(nphotons, 1))
        cta = self._buffer('filter_cta_counts', 4*grid)
        counters = np.zeros(2, dtype=np.uint32)          # dropped, accepted
        cbuf = self.cl_r_buffer('filter_counters', counters)
        ibuf, fbuf = self._cl_buffers['rw_int'], self._cl_buffers['rw_float']
        ev0, ev1 = self._events
        ev0.record(self._stream)
        mod.kernel('TraceFilterFlags').launch(self._stream, grid, block, [
            np.uint32(nphotons), tp, fcfg, rbuf, ibuf, fbuf, flags, cta, cbuf])
        mod.kernel('TraceFilterScan').launch(self._stream, 1, 1024, [
            np.uint32(grid), cta, (cbuf, 0)])
        cbuf.download(self._stream, counters)
        n_dropped, n_sel = int(counters[0]), int(counters[1])
        maxlen = int(trace.maxlen)
        # (the accepted count varies from run to run: grow with 50 % headroom so
        # that a few more rows do not cost a cuMemFree + cuMemAlloc every run)
        obuf_i = self._buffer_with_headroom('trace_compact_int', 4*max(n_sel, 1))
        obuf_f = self._buffer_with_headroom('trace_compact_float', 32*maxlen*max(n_sel, 1))
        mod.kernel('TraceCompact').launch(self._stream, grid, block, [
            np.uint32(nphotons), tp, flags, cta, ibuf, fbuf, obuf_i, obuf_f])
        ev1.record(self._stream)
        self._stream.synchronize()
        self._run_report['filter_ms'] = ev0.elapsed_ms(ev1)
        self._run_report['filter_accepted'] = n_sel
        n_host = np.zeros((n_sel,), dtype=self._types.np_int)
        # (TraceCompact writes whole rows including their zero tails: no zero fill)
        rows = np.empty((n_sel, maxlen) if download else (0, maxlen), dtype=trace.dtype(self))
        if n_sel:
            obuf_i.download(self._stream, n_host)
            if download:
                obuf_f.download(self._stream, rows)

        self._device_trace = dict(n=n_sel, maxlen=maxlen, ints=obuf_i, floats=obuf_f)
        return n_host, rows, n_dropped

    # -- sampling volume (config 4) ----------------------------------------------
    # True: the integer grid of a SamplingVolume object stays on the device across
    # ``sampling_volume`` calls (exact 64-bit sums) and is converted / downloaded when
    # the object's ``data`` is read; False: the reference's flow, one conversion and
    # one download of the whole grid per call (mc.py:1183-1213).
    lazy_sampling_volume = False
    _sv_resident = None              # (weakref to the SamplingVolume, grid cells)

    def _sv_resident_buffer(self, sv, cells: int):
        """Device grid that accumulates for ``sv``; another object's pending grid is
        collected first, a fresh grid starts at zero."""
        cur = self._sv_resident
        owner = cur[0]() if cur is not None else None
        if owner is not sv or cur[1] != cells:
            if owner is not None:
                owner._materialize()
            self._sv_resident = None
        buf = self._buffer('sv_resident', cells*8)
        if self._sv_resident is None:
            buf.fill(self._stream, 0, np.uint32, count=cells*2)
            self._sv_resident = (weakref.ref(sv), cells)
        return buf

    def _sv_loader(self, cells: int):
        """Loader handed to the SamplingVolume: converts the resident grid on the
        device (AccuScale), copies it to the host and restarts the grid at zero."""
        def load(sv):
            cur = self._sv_resident
            if cur is None or cur[0]() is not sv:
                return
            buf = self._cl_buffers['sv_resident']
            whole = type('A', (), {'offset': 0, 'size': cells})
            # (a fresh grid takes a free page-locked buffer of the result pool as it is -
            # no second 64 MB host copy)
            inv_k = 1.0/(sv.k*sv._multiplier(self))
            if sv._data is None:
                scaled, owned = self._scale_on_device(buf, whole, inv_k, want_owned=True)
            else:
                scaled, owned = self._scale_on_device(buf, whole, inv_k), False
            if sv._data is not None:
                sv._data += np.reshape(scaled, sv.shape)
            elif owned:
                sv._data = scaled.reshape(sv.shape)
            else:
                sv._data = np.array(scaled, dtype=np.float64).reshape(sv.shape)
            self._sv_resident = None
        return load

    _SV_SRC = '#define XO_DOUBLE {dbl}\n#define XO_DETERMINISTIC {det}\n#include "xo_sv_kernel.cuh"\n'

    def _pack_sampling_volume(self, trace, sv, nphotons: int):
        """Clears the allocators and packs the trace rows + the sampling volume
        the way the reference does before its ``SamplingVolume`` kernel
        (mc.py:1098-1107).  Returns (packed McTrace, packed McSamplingVolume)."""
        for a in self._allocators.values():
            a.clear()
        self._packed['sv_trace'] = trace.cl_pack(
            self, self._packed.get('sv_trace'), nphotons=int(nphotons))
        self._packed['sv'] = sv.cl_pack(self, self._packed.get('sv'))
        return self._packed['sv_trace'], self._packed['sv']

    sv_warp_per_packet = True        # throughput-mode SamplingVolume kernel (developer knob)

    def sampling_volume(self, trace, sv, wgsize: int = None, maxthreads: int = None,
                        exportsrc: str = None, verbose: bool = False,
                        download: bool = True):
        """Accumulate the sampling volume ``sv`` from the packets of ``trace``
        (counterpart of ``Mc.sampling_volume``, mc.py:1040-1215)."""
        if self._trace is None:
            raise RuntimeError(
                'This Monte Carlo simulator was build without the Trace '
                'functionallity (the trace argument of :py:meth:Mc.__init__) '
                'was None! Sampling volume analysis requires trace functionality!')
        t0 = time.perf_counter()
        self._ensure_device()
        if trace is None:
            # the rows of the last run (or of filter_trace_on_device), in place
            dev = getattr(self, '_device_trace', None)
            if dev is None:
                raise RuntimeError('No trace rows are resident on the device!')
            trace = _ResidentRows(dev, self._trace)
        nphotons = int(trace.nphotons)
        deterministic = self.deterministic
        src = self._SV_SRC.format(det=int(deterministic), dbl=int(bool(
            self.resolved_options().get('MC_USE_DOUBLE_PRECISION', False))))
        if exportsrc:
            with open(exportsrc, 'w') as f:
                f.write(src)
        # throughput mode, single precision: one warp per packet, one lane per segment
        # (xo_sv_kernel.cuh); the reference-structured kernel otherwise
        warp_kernel = bool(self.sv_warp_per_packet and not deterministic and
                           np.dtype(self._types.np_float).itemsize == 4)
        kernel = self._module(src, deterministic).kernel(
            'SamplingVolumeWarp' if warp_kernel else 'SamplingVolume')
        tp, sp = self._pack_sampling_volume(trace, sv, nphotons)
        counters = np.zeros(4, dtype=np.uint32)   # processed, kernels, steps (u64)
        cbuf = self.cl_r_buffer('counters', counters)
        total = np.zeros(1, dtype=np.uint64)
        tbuf = self.cl_r_buffer('sv_total_weight', total)
        sv_allocs = self.cl_rw_accumulator_allocator.allocations(sv)
        lazy = bool(self.lazy_sampling_volume and download and len(sv_allocs) == 1 and
                    int(sv_allocs[0].offset) == 0 and hasattr(sv, '_set_pending') and
                    type(sv).update_data is mcsv_module.SamplingVolume.update_data)
        if lazy:
            # the kernel adds onto the grid that stays on the device for `sv`
            cells = int(sv_allocs[0].size)
            abuf = self._sv_resident_buffer(sv, cells)
        else:
            abuf = self._rw_flat_buffer('accumulator')
        itemsize = np.dtype(self._types.np_float).itemsize
        dev = getattr(self, '_device_trace', None)
        resident = bool(
            nphotons and dev is not None and dev.get('token') is not None and
            dev.get('token') is getattr(trace, '_device_token', None) and
            dev['n'] == nphotons and dev['maxlen'] == int(trace.maxlen))
        if resident:
            # rows never left the device: read them where the run / the device
            # filter put them (offsets folded into the pointers)
            ibuf_arg, fbuf_arg = dev['ints'], dev['floats']
            tp_arg = type(tp)()
            ctypes.memmove(ctypes.addressof(tp_arg), ctypes.addressof(tp), ctypes.sizeof(tp))
            tp_arg.data_buffer_offset = 0
            tp_arg.count_buffer_offset = 0
        else:
            tp_arg = tp
            fbuf = self._rw_flat_buffer('float', fill=False)
            ibuf = self._rw_flat_buffer('int', fill=False)
            ibuf_arg, fbuf_arg = ibuf, fbuf
            self._device_trace = None      # the flat buffers are overwritten
            if nphotons:
                n_host = np.ascontiguousarray(trace.n, dtype=self._types.np_int)
                d_host = np.ascontiguousarray(trace.data, dtype=trace.dtype(self)).view(
                    self._types.np_float).reshape(-1)
                ibuf.upload(self._stream, n_host, offset=int(tp.count_buffer_offset)*4,
                            blocking=True)
                fbuf.upload(self._stream, d_host,
                            offset=int(tp.data_buffer_offset)*itemsize, blocking=True)
        block = int(wgsize) if wgsize else 256
        grid, block = self.launch_geometry(kernel, block, 0, maxthreads)
        if nphotons:
            per_cta = block//32 if warp_kernel else block
            grid = max(1, min(grid, (nphotons + per_cta - 1)//per_cta))
        t1 = time.perf_counter()
        ev0, ev1 = self._events
        ev0.record(self._stream)
        kernel.launch(self._stream, grid, block, [
            np.uint32(nphotons), (cbuf, 0), (cbuf, 4), tp_arg, sp, tbuf, ibuf_arg, fbuf_arg,
            abuf])
        ev1.record(self._stream)
        self._stream.synchronize()
        t2 = time.perf_counter()
        cbuf.download(self._stream, counters)
        tbuf.download(self._stream, total)
        accus = []
        allocs = sv_allocs if (download and not lazy) else ()
        if lazy:
            sv.add_weight(total[0])
            sv._set_pending(self._sv_loader(cells))
        if len(allocs) == 1 and allocs[0].size >= self.SCALE_ON_DEVICE_MIN and \
                type(sv).update_data is mcsv_module.SamplingVolume.update_data:
            # large grid: the float64 conversion of update_data runs on the device
            # (AccuScale, bit-identical), the host copies the page-locked result
            scaled = self._scale_on_device(abuf, allocs[0], 1.0/(sv.k*sv._multiplier(self)))
            sv.update_scaled(scaled, total_weight=total[0])
        else:
            for a in allocs:
                host = self._download_host_array('accumulator', a)
                abuf.download(self._stream, host, offset=a.offset*8)
                accus.append(host)
        if accus:
            sv.update_data(self, accumulators=accus, nphotons=nphotons,
                           total_weight=total[0])
        self._run_report.update(
            upload=t1 - t0, execution=t2 - t1, download=time.perf_counter() - t2,
            items=nphotons, threads=int(counters[1]), sv_kernel_ms=ev0.elapsed_ms(ev1),
            sv_rows_resident=resident,
            sv_steps=int(counters[2:4].view(np.uint64)[0]), sv_grid=grid, sv_block=block)
        if verbose:
            print('SamplingVolume processed {:d} packets in {:d} threads: kernel '
                  '{:.3f} ms'.format(nphotons, int(counters[1]),
                                     self._run_report['sv_kernel_ms']))
        return sv

    # -- raw access (tests / multi-GPU reduction) ---------------------------------
    def download_raw(self):
        """Flat (accumulators, ints, floats) buffers of the last run."""
        out = []
        for kind in ('accumulator', 'int', 'float'):
            alloc = self._allocators[kind]
            host = np.zeros(max(alloc.size, 1), dtype=alloc.dtype)
            self._cl_buffers['rw_' + kind].download(self._stream, host)
            out.append(host)
        return tuple(out)

    def download_seeds(self) -> np.ndarray:
        x = np.empty_like(self._rng_seeds_x)
        self._cl_buffers['rng_seeds_x'].download(self._stream, x)
        return x

    def rng_test(self, n: int, x=None, a=None) -> np.ndarray:
        """Device draw sequence of one (x, a) pair (``RngKernel`` hook,
        mc.py:1320-1362): known-answer test for seed compatibility."""
        from .rngkernel import rng_test
        x = self._rng_seeds_x[0] if x is None else x
        a = self._rng_seeds_a[0] if a is None else a
        return rng_test(self, int(n), int(x), int(a))
