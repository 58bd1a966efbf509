"""Detectors of the cylindrical simulator (mirror of ``xopto/mccyl/mcdetector``:
Detectors container {outer, specular}, FiZ, Total)."""
import numpy as np

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil.axis import Axis, RadialAxis  # noqa: F401
from ..mcml.mcdetector import DetectorBase, Detector, DetectorDefault, NONE

OUTER, SPECULAR = 'outer', 'specular'


def _set_location(det, location):
    # the location vocabulary of this geometry is {outer, specular}
    if location != det._location and det._location != NONE:
        raise RuntimeError('Detector location cannot be changed!')
    det._location = location


class FiZ(Detector):
    """Azimuth-z grid on the outer sample surface; raw data indexed [z, fi]
    (mccyl/mcdetector/fiz.py)."""
    cu_type = 'xo::DetFiZ'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFiZ(cltypes.Structure):
            _pack_ = 1
            _fields_ = [('fi_min', T.mc_fp_t), ('inv_dfi', T.mc_fp_t),
                        ('z_min', T.mc_fp_t), ('inv_dz', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n_fi', T.mc_size_t),
                        ('n_z', T.mc_size_t), ('offset', T.mc_size_t)]
        return ClFiZ

    def __init__(self, fiaxis, zaxis=None, cosmin: float = 0.0):
        if isinstance(fiaxis, FiZ):
            o = fiaxis
            fiaxis, zaxis = type(o.fiaxis)(o.fiaxis), type(o.zaxis)(o.zaxis)
            cosmin = o.cosmin
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if zaxis is None:
                zaxis = Axis(-1.0, 1.0, 1)
            raw, nphotons = np.zeros((zaxis.n, fiaxis.n)), 0
        super().__init__(raw, nphotons)
        self._fi_axis, self._z_axis = fiaxis, zaxis
        self.cosmin = cosmin
        self._r_sample = 1.0
        self._accumulators_area = fiaxis.step*zaxis.step

    fiaxis = property(lambda self: self._fi_axis)
    zaxis = property(lambda self: self._z_axis)
    fi = property(lambda self: self._fi_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    fiedges = property(lambda self: self._fi_axis.edges)
    zedges = property(lambda self: self._z_axis.edges)
    nfi = property(lambda self: self._fi_axis.n)
    nz = property(lambda self: self._z_axis.n)

    def meshgrid(self):
        return np.meshgrid(self.z, self.fi, indexing='ij')

    def update_data(self, mc, *args, **kwargs):
        self._r_sample = mc.layers[1].d*0.5
        return super().update_data(mc, *args, **kwargs)

    @property
    def normalized(self):
        area = self._accumulators_area*self._r_sample
        return self.raw*(1.0/(max(self.nphotons, 1.0)*area))

    reflectance = property(lambda self: self.normalized)
    transmittance = property(lambda self: self.normalized)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.fi_min = self._fi_axis.start
        target.inv_dfi = 1.0/self._fi_axis.step if self._fi_axis.n > 1 else 0.0
        target.z_min = self._z_axis.start
        target.inv_dz = 1.0/self._z_axis.step if self._z_axis.n > 1 else 0.0
        target.cos_min = self.cosmin
        target.n_fi, target.n_z = self._fi_axis.n, self._z_axis.n
        return target

    def todict(self):
        return {'type': 'FiZ', 'fi_axis': self._fi_axis.todict(),
                'z_axis': self._z_axis.todict(), 'cosmin': self._cosmin}


class Total(Detector):
    """Single accumulator; acceptance measured against the radial surface
    normal (mccyl/mcdetector/total.py)."""
    cu_type = 'xo::DetTotalCyl'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClTotal(cltypes.Structure):
            _fields_ = [('cos_min', T.mc_fp_t), ('offset', T.mc_size_t)]
        return ClTotal

    def __init__(self, cosmin: float = 0.0):
        if isinstance(cosmin, Total):
            o = cosmin
            cosmin, raw, nphotons = o.cosmin, np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((1,)), 0
        super().__init__(raw, nphotons)
        self.cosmin = cosmin

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.cos_min = self._cosmin
        return target

    def todict(self):
        return {'type': 'Total', 'cosmin': self._cosmin}


class Detectors(McObject):
    """Container {outer, specular} (mccyl/mcdetector/base.py:280-520)."""

    def __init__(self, outer=None, specular=None):
        super().__init__()
        if isinstance(outer, Detectors):
            d = outer
            outer, specular = type(d.outer)(d.outer), type(d.specular)(d.specular)
        outer = DetectorDefault() if outer is None else outer
        specular = DetectorDefault() if specular is None else specular
        if isinstance(outer, DetectorDefault) and outer.location != NONE:
            outer = DetectorDefault()
        if isinstance(specular, DetectorDefault) and specular.location != NONE:
            specular = DetectorDefault()
        _set_location(outer, OUTER)
        _set_location(specular, SPECULAR)
        self._outer, self._specular = outer, specular

    outer = property(lambda self: self._outer)
    specular = property(lambda self: self._specular)

    def cl_type(self, mc):
        class ClDetectors(cltypes.Structure):
            _fields_ = [('outer', self._outer.fetch_cl_type(mc)),
                        ('specular', self._specular.fetch_cl_type(mc))]
        return ClDetectors

    def cl_options(self, mc):
        options, used = [], False
        for det, name in ((self._outer, 'OUTER'), (self._specular, 'SPECULAR')):
            if type(det) is not DetectorDefault:
                options.append(('MC_USE_{}_DETECTOR'.format(name), True))
                options.extend(det.fetch_cl_options(mc))
                used = True
        if used:
            options.append(('MC_USE_DETECTORS', True))
        return options

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        self._outer.cl_pack(mc, target.outer)
        self._specular.cl_pack(mc, target.specular)
        return target

    def update_data(self, mc, detector, data, nphotons=0):
        location = detector.location if isinstance(detector, DetectorBase) else detector
        if location not in (OUTER, SPECULAR):
            raise ValueError('Detector location must be one of "{}" or "{}" '
                             'but got "{}"!'.format(OUTER, SPECULAR, location))
        getattr(self, location).update_data(
            mc, accumulators=data.get(np.dtype(mc.types.np_accu)),
            float_buffers=data.get(np.dtype(mc.types.np_float)),
            integer_buffers=data.get(np.dtype(mc.types.np_int)), nphotons=nphotons)

    def types(self):
        return type(self._outer), type(self._specular)

    def __iter__(self):
        return iter([self._outer, self._specular])

    def todict(self):
        return {'type': 'Detectors', 'outer': self._outer.todict(),
                'specular': self._specular.todict()}
