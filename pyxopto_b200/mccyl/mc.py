"""Concentric-cylinder simulator ``Mc`` (drop-in for ``xopto.mccyl.mc.Mc``,
mccyl/mc.py:75-1015) on top of the CUDA kernel ``csrc/kernels/mccyl_kernel.cuh``.

    from pyxopto_b200.mccyl import mc
    layers = mc.mclayer.Layers([
        mc.mclayer.Layer(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),     # surrounding
        mc.mclayer.Layer(d=10e-3, n=1.33, mua=1e2, mus=100e2, pf=pf)])
    det = mc.mcdetector.Detectors(outer=mc.mcdetector.FiZ(fiaxis, zaxis))
    sim = mc.Mc(layers, mc.mcsource.Line((-10e-3, 0, 0)), det)
    trace, fluence, detectors = sim.run(1e6)
"""
import ctypes

import numpy as np

from ..cl import clinfo, clrng, cltypes            # noqa: F401
from ..mcbase import mcoptions, mctypes, mcobject  # noqa: F401
from ..mcbase import mcsv, mcprogress                         # noqa: F401
from ..mcbase import mcpf, mcfluence, mctrace      # noqa: F401
from ..mcbase.mcobject import McObject             # noqa: F401
from ..mcbase.mcsim import McBase
from . import mclayer, mcsource, mcdetector        # noqa: F401


class Mc(McBase):
    kernel_header = 'mccyl_kernel.cuh'
    geometry = 'mccyl'
    default_refill_lanes = 4         # waiting lanes per warp that trigger a service round

    def __init__(self, layers, source, detectors=None, trace=None, fluence=None,
                 surface=None, types=mctypes.McDataTypesSingle, options=None,
                 rnginit=None, cl_devices=None, cl_build_options=None,
                 cl_profiling: bool = False):
        super().__init__(source, detectors=detectors, trace=trace, fluence=fluence,
                         surface=surface, types=types, options=options,
                         rnginit=rnginit, cl_devices=cl_devices,
                         cl_build_options=cl_build_options, cl_profiling=cl_profiling)
        from ..adopt import adopt
        layers = adopt(layers, self.geometry)
        if not isinstance(layers, mclayer.Layers):
            layers = mclayer.Layers([adopt(item, self.geometry) for item in layers])
        self._layers = layers
        self._obj_types['layer'] = type(layers[1])
        self._obj_types['pf'] = type(layers[1].pf)

    layers = property(lambda self: self._layers)

    def layer(self, index: int):
        return self._layers[index]

    def layer_index(self, r: float) -> int:
        return self._layers.layer_index(r)

    # -- packing -----------------------------------------------------------------
    user_plugin_slots = ('XoPf', 'XoSource', 'XoDetOuter', 'XoDetSpecular', 'XoFluence',
                         'XoTrace')
    clcompat_geometry_header = 'xo_clcompat_mccyl.cuh'

    def _plugin_objects(self):
        dets = self._detectors
        return {'XoPf': self._layers[1].pf, 'XoSource': self._source,
                'XoDetOuter': dets.outer if dets is not None else None,
                'XoDetSpecular': dets.specular if dets is not None else None,
                'XoFluence': self._fluence}

    def _scattering_pfs(self):
        return [item.pf for item in list(self._layers)[1:]]

    def _pack_medium(self):
        if type(self._layers[1].pf) is not self._obj_types['pf']:
            raise ValueError('The scattering phase function kind/type must not '
                             'change between simulation calls!')
        self._packed['layers'] = self._layers.cl_pack(self, self._packed.get('layers'))

    def _medium_bytes(self) -> int:
        # packed layers + the per-layer derived constants the throughput kernel
        # appends (xo::CylFastLayer, <= 112 B per layer)
        return len(cltypes.raw_bytes(self._packed['layers'])) + 112*len(self._layers) + 32

    def _upload_medium(self):
        self.cl_r_buffer('layers', self._packed['layers'])

    # -- translation unit ----------------------------------------------------------
    def _detector_bindings(self):
        dets = self._detectors
        out = []
        for loc, Name in (('outer', 'XoDetOuter'), ('specular', 'XoDetSpecular')):
            det = getattr(dets, loc) if dets is not None else mcdetector.DetectorDefault()
            out.append((Name, det.fetch_cu_type(self), det.fetch_cl_type(self)))
        return out

    def _extra_defines(self, opts):
        aniso = isinstance(self._layers[1], mclayer.AnisotropicLayer)
        return ['#define XO_ANISO {}'.format(int(aniso))]

    def _plugin_bindings(self):
        pf = self._layers[1].pf
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
        return ['#include "mccyl_sources.cuh"']

    def _extra_checks(self):
        checks = ['static_assert(sizeof(xo::CylLayer) == {}, "McLayer layout differs '
                  'from the packed host struct");'.format(
                      ctypes.sizeof(self._layers[0].fetch_cl_type(self)))]
        if self._detectors is not None:
            checks.append('static_assert(sizeof(xo::XoDetectors) == {}, "McDetectors '
                          'layout differs from the packed host struct");'.format(
                              ctypes.sizeof(self._detectors.fetch_cl_type(self))))
        return checks

    def kernel_source(self, block=256, min_blocks=1):
        opts = self.resolved_options()
        if int(opts.get('MC_METHOD', 0)) == 2:
            # the reference's MBL branch does not compile for this geometry
            # (undeclared d_ok, mccyl.template.c:774,781)
            raise NotImplementedError(
                'The microscopic Beer-Lambert method is not available in mccyl.')
        return super().kernel_source(block=block, min_blocks=min_blocks)

    # -- launch ---------------------------------------------------------------------
    def _kernel_args(self, nphotons, bufs, lut_len, priv_len, chunk, refill, window):
        if self._detectors is not None:
            dets = self._packed['detectors']
        else:
            dets = mcdetector.Detectors().cl_pack(self)
        return [
            np.uint32(nphotons),
            (bufs['counters'], 0),            # num_packets_done
            (bufs['counters'], 4),            # num_kernels
            self._types.np_float(self._rmax),
            bufs['rng_x'], bufs['rng_a'],
            np.uint32(len(self._layers)),
            self._cl_buffers['layers'],
            self._packed['source'],
            self._packed_or_dummy('trace', 4),
            self._packed_or_dummy('fluence', 4),
            dets,
            bufs['lut'], bufs['ints'], bufs['floats'], bufs['accu'],
            np.uint32(lut_len), np.uint32(priv_len), window, np.uint32(max(chunk, 1)),
            np.uint32(max(refill, 1)),
        ]
