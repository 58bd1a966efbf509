"""Voxelised simulator ``Mc`` (drop-in for ``xopto.mcvox.mc.Mc``, mcvox/mc.py:76-1100)
on top of the CUDA kernel ``csrc/kernels/mcvox_kernel.cuh``.

    from pyxopto_b200.mcvox import mc
    voxels = mc.mcgeometry.Voxels(xaxis, yaxis, zaxis)
    sim = mc.Mc(voxels, materials, mc.mcsource.GaussianBeam(50e-6), fluence=...)
    sim.voxels.material[...] = 1
    trace, fluence, detectors = sim.run(1e6)
"""
import ctypes

import os

import numpy as np

from ..cl import clinfo, clrng, cltypes            # noqa: F401
from ..mcbase import mctypes, mcobject             # noqa: F401
from . import mcoptions                            # noqa: F401  (shared + McMaterialMemory)
from ..mcbase import mcsv, mcprogress                         # noqa: F401
from ..mcbase import mcpf, mcfluence, mctrace, mcmaterial  # noqa: F401
from ..mcbase.mcobject import McObject             # noqa: F401
from ..mcbase.mcsim import McBase
from ..mcml import mcdetector                      # noqa: F401  (same detector family)
from . import mcgeometry, mcsource                 # noqa: F401


class Mc(McBase):
    kernel_header = 'mcvox_kernel.cuh'
    fluence_block = 1024
    # no CTA-private fluence window: deposits are spread over the whole grid (a
    # window around the beam catches few of them) and the shared memory is worth
    # more as L1 cache of the compact voxel map (measured: 614 vs 465 Mpackets/s)
    fluence_window_bytes = 0
    geometry = 'mcvox'

    def __init__(self, voxels, materials, source, detectors=None, trace=None,
                 fluence=None, surface=None, types=mctypes.McDataTypesSingle,
                 options=None, rnginit=None, cl_devices=None, cl_build_options=None,
                 cl_profiling: bool = False):
        super().__init__(source, detectors=detectors, trace=trace, fluence=fluence,
                         surface=surface, types=types, options=options,
                         rnginit=rnginit, cl_devices=cl_devices,
                         cl_build_options=cl_build_options, cl_profiling=cl_profiling)
        from ..adopt import adopt
        voxels, materials = adopt(voxels, self.geometry), adopt(materials, self.geometry)
        if not isinstance(materials, mcmaterial.Materials):
            materials = mcmaterial.Materials([adopt(item, self.geometry) for item in materials])
        self._voxels = voxels
        self._materials = materials
        self._voxels.data(self)            # allocate the material-index array
        self._obj_types['pf'] = type(materials[0].pf)
        self._voxels_on_device = False

    voxels = property(lambda self: self._voxels)
    materials = property(lambda self: self._materials)

    def material(self, index):
        """Material by index or at a position (mcvox/mc.py:420-450).  The
        reference indexes the [z, y, x] array with an (x, y, z) tuple here; kept
        for parity (it only matters for the specular term of asymmetric grids)."""
        material_index = 0
        if isinstance(index, (int, np.integer)):
            material_index = int(index)
        elif self._voxels.contains(index):
            ind = self._voxels.index(index)
            material_index = int(self._voxels[ind]['material_index'])
        return self._materials[material_index]

    # -- packing -----------------------------------------------------------------
    user_plugin_slots = ('XoPf', 'XoSource', 'XoDetTop', 'XoDetBottom', 'XoDetSpecular',
                         'XoFluence', 'XoTrace')
    clcompat_geometry_header = 'xo_clcompat_mcvox.cuh'

    def _plugin_objects(self):
        dets = self._detectors
        return {'XoPf': self._materials[0].pf, 'XoSource': self._source,
                'XoDetTop': dets.top if dets is not None else None,
                'XoDetBottom': dets.bottom if dets is not None else None,
                'XoDetSpecular': dets.specular if dets is not None else None,
                'XoFluence': self._fluence}

    def _scattering_pfs(self):
        return [item.pf for item in list(self._materials)]

    def _pack_medium(self):
        if type(self._materials[0].pf) is not self._obj_types['pf']:
            raise ValueError('The scattering phase function kind/type must not '
                             'change between simulation calls!')
        self._packed['materials'] = self._materials.cl_pack(
            self, self._packed.get('materials'))
        self._packed['voxels'] = self._voxels.cl_pack(self, self._packed.get('voxels'))

    def _medium_bytes(self) -> int:
        # packed materials + the per-material derived constants the throughput
        # loop appends (xo::VoxFastMat, <= 64 B per material)
        return len(cltypes.raw_bytes(self._packed['materials'])) + \
            64*len(self._materials) + 32

    # waiting lanes per warp that trigger their joint handling (throughput loop).
    # Measured on C3 (201^3, ms per 1e8 packets) with the clearance map, where 68 % of
    # the flights skip the voxel walk and wait for their interaction right away:
    # 16: 121.2, 20: 113.4, 22: 112.2, 24: 112.8, 26: 116.4, 28: 124.3, 32: 187.1
    # (before the clearance map the optimum was 16: 129.6)
    wait_lanes = 22

    def _refill_lanes(self) -> int:
        if self.refill_lanes is not None:
            return int(min(max(self.refill_lanes, 1), 32))
        return int(self.wait_lanes)

    def _rmax_needed(self) -> bool:
        """Packets only exist inside the voxel box: the rmax test can fire only
        if some corner of the box is farther than rmax from the source."""
        rmax = float(np.float32(self._rmax))
        if not np.isfinite(rmax):
            return False
        src = np.asarray(self._source_focus(), dtype=np.float64)
        v = self._voxels
        far = 0.0
        for x in (v.xaxis.start, v.xaxis.stop):
            for y in (v.yaxis.start, v.yaxis.stop):
                for z in (v.zaxis.start, v.zaxis.stop):
                    far = max(far, float(np.linalg.norm(np.array([x, y, z]) - src)))
        return far*(1.0 + 1e-5) >= rmax

    # -- compact voxel map of the throughput loop -------------------------------------
    # uint8 material indices in a box padded by two voxels of the sentinel 255 on
    # every side (the walk looks one voxel ahead speculatively), axis strides
    # rounded up to powers of two: the linear index IS the packed coordinate triple
    # (x+2) | (y+2) << bx | (z+2) << (bx+by), and leaving the grid is just another
    # "material change".
    VOX_SENTINEL = 255

    def _vox_pack_bits(self):
        nz, ny, nx = self._voxels.shape
        return tuple(int(n + 3).bit_length() for n in (nx, ny, nz))

    @staticmethod
    def _clearance(mat: np.ndarray) -> np.ndarray:
        """Chessboard distance (voxels, 1 ... 255) of every voxel to the nearest voxel
        of another material or outside the grid: all voxels closer than that - in the
        maximum norm - hold the same material.  The throughput loop lets a flight that
        is shorter than the clearance skip the voxel walk (mcvox_dda_loop.cuh)."""
        from scipy import ndimage
        padded = np.full(tuple(n + 2 for n in mat.shape), -1, dtype=np.int32)
        padded[1:-1, 1:-1, 1:-1] = mat
        dist = np.zeros(padded.shape, dtype=np.int32)
        for m in np.unique(mat):
            dist += ndimage.distance_transform_cdt(padded == m, metric='chessboard')
        return np.clip(dist[1:-1, 1:-1, 1:-1], 1, 255).astype(np.uint8)

    def _vox_packed(self) -> bool:
        return len(self._materials) <= self.VOX_SENTINEL and sum(self._vox_pack_bits()) <= 31

    def _same_refractive_index(self) -> bool:
        """True when every material (the surrounding medium included) packs the same
        refractive index: a face between two materials then never reflects or
        refracts, and the throughput loop treats it as a change of the attenuation
        along the ray instead of an interface event."""
        ns = {float(np.float32(m.n)) for m in self._materials}
        return len(ns) == 1

    def _fluence_on_voxel_grid(self) -> bool:
        """True when the fluence plugin is the plain ``Fluence`` accumulator on exactly
        the voxel grid: the cell of a deposit is then the voxel the walk is in."""
        flu = self._fluence
        if flu is None or type(flu) is not mcfluence.Fluence:
            return False
        v = self._voxels
        for fa, va in ((flu.xaxis, v.xaxis), (flu.yaxis, v.yaxis), (flu.zaxis, v.zaxis)):
            if (np.float32(fa.start), np.float32(fa.stop), int(fa.n)) != \
                    (np.float32(va.start), np.float32(va.stop), int(va.n)) or \
                    getattr(fa, 'logscale', False):
                return False
        return True

    # Per-warp packet pool of the throughput loop (csrc/kernels/mcvox_pool_loop.cuh): a warp
    # owns 64 packet slots in shared memory and runs, every round, the phase most of them
    # wait for.  0 keeps the packets in the registers of their lane (mcvox_dda_loop.cuh).
    pool_slots = 64
    pool_tuning = {}                 # XO_POOL_* thresholds (developer knob)
    # A full trace (one event per loop trip) is a store stream: the SM's store path, not the
    # lane count, bounds it, and the lane-resident loop with its 10 walking lanes per warp
    # feeds it better than the pool with 25 (C4 on the voxel geometry, 2e5 packets: 1.24 ms
    # against 1.88 ms; more resident warps lose the same way, profiles/r02y_*).  The pool
    # records traces correctly (tests) and is kept for start / end traces.
    pool_full_trace = False
    # rings of slot numbers per class (default) or a census of the slot states + a gather by
    # rank every round (C3: 78.7 against 80.7 ms per 1e8 packets, profiles/r03r_*)
    pool_queues = bool(int(os.environ.get('XOPTO_POOL_QUEUES', '1')))

    def _pool_slots(self, opts=None) -> int:
        """Slots per warp, or 0 where the pool loop does not apply: it covers the compact
        map in throughput mode with albedo weight / rejection and isotropic materials
        (start / end traces included; a full trace keeps the lane-resident loop unless
        ``pool_full_trace``, a user-written trace always)."""
        opts = self.resolved_options() if opts is None else opts
        if not self.pool_slots or self.deterministic or not self._vox_packed() or \
                self._user_trace() or \
                (int(opts.get('MC_USE_TRACE', 0)) == 7 and not self.pool_full_trace) or \
                isinstance(self._materials[0], mcmaterial.AnisotropicMaterial) or \
                opts.get('MC_METHOD', 0) not in (0, 1):
            return 0
        return 64                    # (the census of the loop reads two slots per lane)

    def _loop_name(self) -> str:
        opts = self.resolved_options()
        if self.deterministic or not self._vox_packed() or opts.get('MC_METHOD', 0) == 2:
            return 'reference-structured'
        return 'packet pool' if self._pool_slots(opts) else 'lane-resident rays'

    def _queue_bytes(self, block: int) -> int:
        slots = self._pool_slots()
        if slots:
            # 4 x float4 + 1 float (traced packets: a fifth float4) + 1 state byte per slot,
            # 32 index bytes per warp
            # ... where the rmax sphere can be reached the ray parameter of its exit (float)
            per_slot = 64 + (16 if self._trace is not None else 4) + \
                (4 if self._rmax_needed() else 0) + 1
            # ... and five rings of 64 slot numbers per warp (XO_POOL_QUEUES)
            return (block//32)*(slots*per_slot + 32 + 320) + 32
        return super()._queue_bytes(block)

    def _extra_defines(self, opts):
        aniso = isinstance(self._materials[0], mcmaterial.AnisotropicMaterial)
        return ['#define XO_VOX_POOL {}'.format(self._pool_slots(opts)),
                '#define XO_POOL_QUEUES {}'.format(int(bool(self.pool_queues)))] + \
            ['#define {} {}'.format(k, int(v)) for k, v in sorted(self.pool_tuning.items())] + \
            ['#define XO_VOX_PACKED {}'.format(int(self._vox_packed())),
                '#define XO_ANISO {}'.format(int(aniso)),
                '#define XO_VOX_SAME_N {}'.format(
                    int(self._same_refractive_index() and not aniso)),
                '#define XO_FLU_VOXGRID {}'.format(int(self._fluence_on_voxel_grid()))]

    def _upload_medium(self):
        self.cl_r_buffer('materials', self._packed['materials'])
        if self._voxels.update_required() or not self._voxels_on_device:
            data = np.ascontiguousarray(self._voxels.data(self)).view(np.int32)
            self.cl_r_buffer('voxel_data', data)
            if self._vox_packed():
                bx, by, bz = self._vox_pack_bits()
                nz, ny, nx = self._voxels.shape
                mat = data.reshape(nz, ny, nx)
                if mat.size and (mat.min() < 0 or mat.max() >= len(self._materials)):
                    raise ValueError('Voxel material indices must be in [0, {})!'.format(
                        len(self._materials)))
                packed = np.full((1 << bz, 1 << by, 1 << bx), self.VOX_SENTINEL, np.uint16)
                packed[2:nz + 2, 2:ny + 2, 2:nx + 2] = \
                    mat.astype(np.uint16) | (self._clearance(mat).astype(np.uint16) << 8)
                # the voxel walk steps the low address word only: the map must not
                # straddle a 4 GB boundary (re-allocate in the unlikely case it does)
                parked = []
                for _ in range(4):
                    buf = self.cl_r_buffer('voxel_packed', packed)
                    if (buf.device_ptr & 0xFFFFFFFF) + packed.nbytes <= 1 << 32:
                        break
                    parked.append(self._cl_buffers.pop('voxel_packed'))
                else:
                    raise RuntimeError('Could not place the compact voxel map inside '
                                       'one 4 GB address window.')
                del parked
            self._voxels_on_device = True

    # -- translation unit ----------------------------------------------------------
    def _plugin_bindings(self):
        pf = self._materials[0].pf
        out = [('XoPf', pf.fetch_cu_type(self), pf.fetch_cl_type(self)),
               ('XoSource', self._source.fetch_cu_type(self),
                self._source.fetch_cl_type(self))]
        out += self._detector_bindings()
        if self._fluence is not None:
            out.append(('XoFluence', self._fluence.fetch_cu_type(self),
                        self._fluence.fetch_cl_type(self)))
        else:
            out.append(('XoFluence', 'xo::FluNone', None))
        return out

    def _extra_includes(self):
        return ['#include "mcvox_sources.cuh"']

    def _extra_checks(self):
        checks = [
            'static_assert(sizeof(xo::VoxMaterial) == {}, "McMaterial layout differs '
            'from the packed host struct");'.format(
                ctypes.sizeof(self._materials[0].fetch_cl_type(self))),
            'static_assert(sizeof(xo::VoxCfg) == {}, "McVoxelConfig layout differs '
            'from the packed host struct");'.format(
                ctypes.sizeof(self._voxels.fetch_cl_type(self)))]
        if self._detectors is not None:
            checks.append('static_assert(sizeof(xo::XoDetectors) == {}, "McDetectors '
                          'layout differs from the packed host struct");'.format(
                              ctypes.sizeof(self._detectors.fetch_cl_type(self))))
        return checks

    # -- launch ---------------------------------------------------------------------
    def _kernel_args(self, nphotons, bufs, lut_len, priv_len, chunk, refill, window):
        if self._detectors is not None:
            dets = self._packed['detectors']
        else:
            dets = mcdetector.Detectors().cl_pack(self)
        packed = self._vox_packed()
        bx, by, _ = self._vox_pack_bits()
        return [
            np.uint32(nphotons),
            (bufs['counters'], 0), (bufs['counters'], 4),
            self._types.np_float(self._rmax),
            bufs['rng_x'], bufs['rng_a'],
            self._packed['voxels'],
            self._cl_buffers['voxel_data'],
            np.uint32(len(self._materials)),
            self._cl_buffers['materials'],
            self._packed['source'],
            self._packed_or_dummy('trace', 4),
            self._packed_or_dummy('fluence', 4),
            dets,
            bufs['lut'], bufs['ints'], bufs['floats'], bufs['accu'],
            np.uint32(lut_len), np.uint32(priv_len), window, np.uint32(max(chunk, 1)),
            np.uint32(refill),
            self._cl_buffers['voxel_packed' if packed else 'voxel_data'],
            np.uint32(bx), np.uint32(by),
        ]
