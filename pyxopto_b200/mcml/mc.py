"""Layered-media simulator ``Mc`` (drop-in for ``xopto.mcml.mc.Mc``,
mcml/mc.py:75-1018) on top of the CUDA kernel ``csrc/kernels/mcml_kernel.cuh``.

Usage mirrors the reference::

    from pyxopto_b200.mcml import mc
    layers = mc.mclayer.Layers([...])
    sim = mc.Mc(layers, mc.mcsource.Line(), mc.mcdetector.Detectors(top=...))
    sim.rmax = 25e-3
    trace, fluence, detectors = sim.run(1e6)
"""
import numpy as np

from ..cl import clinfo, clrng, cltypes            # noqa: F401
from ..mcbase import mcoptions, mctypes, mcobject  # noqa: F401
from ..mcbase import mcsv, mcprogress                         # noqa: F401
from ..mcbase import mcpf, mcfluence, mctrace      # noqa: F401
from ..mcbase.mcobject import McObject             # noqa: F401
from ..mcbase.mcsim import McBase
from . import mclayer, mcsource, mcdetector, mcsurface  # noqa: F401


class Mc(McBase):
    supports_surface_layouts = True
    kernel_header = 'mcml_kernel.cuh'
    geometry = 'mcml'
    # waiting lanes per warp that trigger a service round (measured optimum 2-4
    # for short-launch sources on slabs: C1 9.9e8 packets/s at 3)
    default_refill_lanes = 3

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

    def layer_index(self, z: float) -> int:
        return self._layers.layer_index(z)

    # -- packing -----------------------------------------------------------------
    def _scattering_pfs(self):
        return [item.pf for item in list(self._layers)[1:-1]]

    def _pack_medium(self):
        if type(self._layers[1].pf) is not self._obj_types['pf']:
            raise ValueError('The scattering phase function kind/type must not '
                             'change between simulation calls!')
        self._packed['layers'] = self._layers.cl_pack(self, self._packed.get('layers'))

    # lanes of a warp waiting for a new packet that trigger a pop from the warp's
    # launch queue (throughput mode)
    pop_batch = 1

    def _medium_bytes(self) -> int:
        # packed layers + the per-layer derived constants the kernel appends
        # (xo::MlFastLayer, <= 112 B per layer)
        return len(cltypes.raw_bytes(self._packed['layers'])) + 112*len(self._layers) + 32

    def _upload_medium(self):
        self.cl_r_buffer('layers', self._packed['layers'])

    # -- translation unit ----------------------------------------------------------
    def _extra_defines(self, opts):
        aniso = isinstance(self._layers[1], mclayer.AnisotropicLayer)
        return ['#define XO_ANISO {}'.format(int(aniso))]

    def _plugin_bindings(self):
        pf = self._layers[1].pf
        out = [('XoPf', pf.fetch_cu_type(self), pf.fetch_cl_type(self)),
               ('XoSource', self._source.fetch_cu_type(self),
                self._source.fetch_cl_type(self))]
        out += self._surface_bindings()
        out += self._detector_bindings()
        if self._fluence is not None:
            out.append(('XoFluence', self._fluence.fetch_cu_type(self),
                        self._fluence.fetch_cl_type(self)))
        else:
            out.append(('XoFluence', 'xo::FluNone', None))
        return out

    user_plugin_slots = ('XoPf', 'XoSource', 'XoDetTop', 'XoDetBottom', 'XoDetSpecular',
                         'XoFluence', 'XoSurfTop', 'XoSurfBottom', 'XoTrace')
    clcompat_geometry_header = 'xo_clcompat_mcml.cuh'

    def _plugin_objects(self):
        dets = self._detectors
        return {'XoPf': self._layers[1].pf, 'XoSource': self._source,
                'XoDetTop': dets.top if dets is not None else None,
                'XoDetBottom': dets.bottom if dets is not None else None,
                'XoDetSpecular': dets.specular if dets is not None else None,
                'XoFluence': self._fluence,
                'XoSurfTop': self._surface.top if self._surface is not None else None,
                'XoSurfBottom': self._surface.bottom if self._surface is not None else None}

    def _surface_bindings(self):
        layouts = self._surface if self._surface is not None else mcsurface.SurfaceLayouts()
        return [(name, lay.fetch_cu_type(self), lay.fetch_cl_type(self))
                for name, lay in (('XoSurfTop', layouts.top), ('XoSurfBottom', layouts.bottom))]

    def _extra_includes(self):
        return ['#include "xo_surface.cuh"', '#include "mcml_sources.cuh"']

    def _extra_checks(self):
        import ctypes
        checks = ['static_assert(sizeof(xo::MlLayer) == {}, "McLayer layout differs '
                  'from the packed host struct");'.format(
                      ctypes.sizeof(self._layers[0].fetch_cl_type(self)))]
        if self._surface is not None:
            checks.append('static_assert(sizeof(xo::XoSurface) == {}, "McSurfaceLayouts '
                          'layout differs from the packed host struct");'.format(
                              ctypes.sizeof(self._surface.fetch_cl_type(self))))
        if self._detectors is not None:
            checks.append('static_assert(sizeof(xo::XoDetectors) == {}, "McDetectors '
                          'layout differs from the packed host struct");'.format(
                              ctypes.sizeof(self._detectors.fetch_cl_type(self))))
        return checks

    # -- launch ---------------------------------------------------------------------
    def _kernel_args(self, nphotons, bufs, lut_len, priv_len, chunk, refill, window):
        from . import mcdetector as md
        T = self._types
        if self._detectors is not None:
            dets = self._packed['detectors']
        else:
            dets = md.Detectors().cl_pack(self)
        return [
            T.np_cnt(nphotons) if T.np_cnt is np.uint32 else np.uint32(nphotons),
            (bufs['counters'], 0),            # num_packets_done
            (bufs['counters'], 4),            # num_kernels
            self._types.np_float(self._rmax),
            bufs['rng_x'], bufs['rng_a'],
            np.uint32(len(self._layers)),
            self._cl_buffers['layers'],
            self._packed['source'],
            self._packed['surface_layouts'] if self._surface is not None
            else mcsurface.SurfaceLayouts().cl_pack(self),
            self._packed_or_dummy('trace', 4),
            self._packed_or_dummy('fluence', 4),
            dets,
            bufs['lut'], bufs['ints'], bufs['floats'], bufs['accu'],
            np.uint32(lut_len), np.uint32(priv_len), window,
            np.uint32(min(max(int(self.pop_batch), 1), 32)),
            np.uint32(refill),
        ]
