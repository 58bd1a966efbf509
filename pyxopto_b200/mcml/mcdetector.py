"""Surface / specular detectors (mirror of ``xopto/mcml/mcdetector``: Detectors
container, Total, Radial, Cartesian, SixAroundOne, RadialPl, TotalPl).

Each class packs the reference's ``Mc<Location>Detector`` struct and names the
CUDA struct in ``csrc/kernels/xo_detectors.cuh`` that bins the escaping packets.
"""
from typing import Tuple

import numpy as np

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil import geometry
from ..mcbase.mcutil.axis import Axis, RadialAxis, SymmetricAxis  # noqa: F401
from ..mcbase.mcutil.fiber import MultimodeFiber, FiberLayout  # noqa: F401
from ..mcbase.mcutil.lut import CollectionLut, LinearLut  # noqa: F401

NONE, TOP, BOTTOM, SPECULAR = 'none', 'top', 'bottom', 'specular'


class DetectorBase(McObject):
    def __init__(self, location=NONE):
        super().__init__()
        self._location = location

    def _set_location(self, location):
        if location not in (TOP, BOTTOM, SPECULAR):
            raise ValueError('Detector location must be "{}", "{}", "{}"!'.format(
                TOP, BOTTOM, SPECULAR))
        if location != self._location and self._location != NONE:
            raise RuntimeError('Detector location cannot be changed!')
        self._location = str(location)

    location = property(lambda self: self._location, _set_location)


class DetectorDefault(DetectorBase):
    """Placeholder for an unused location: ``{int64 dummy}`` (base.py:145)."""
    cu_type = 'xo::DetNone'

    def cl_type(self, mc):
        class ClDetectorDefault(cltypes.Structure):
            _fields_ = [('dummy', cltypes.cl_int64_t)]
        return ClDetectorDefault

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.dummy = 0
        return target

    def todict(self):
        return {'type': type(self).__name__}


class Detector(DetectorBase):
    def __init__(self, raw_data: np.ndarray, nphotons: int):
        super().__init__(NONE)
        self._raw_data = raw_data
        self._nphotons = int(nphotons)

    raw = property(lambda self: self._raw_data, None, None, 'Raw accumulator data.')
    nphotons = property(lambda self: self._nphotons)
    shape = property(lambda self: self._raw_data.shape)
    total = property(lambda self: self._raw_data.sum())

    def set_raw_data(self, data, nphotons):
        self._raw_data[:] = np.asarray(data, dtype=self._raw_data.dtype)
        self._nphotons = int(nphotons)

    def update_data(self, mc, accumulators, nphotons, **kwargs):
        new_data = np.reshape(accumulators[0], self.shape)
        self._raw_data += new_data*(1.0/mc.types.mc_accu_k)
        self._nphotons += int(nphotons)

    @property
    def normalized(self):
        return self.raw*(1.0/max(self.nphotons, 1.0))

    reflectance = property(lambda self: self.normalized)
    transmittance = property(lambda self: self.normalized)

    # shared helpers
    def _set_cosmin(self, v):
        self._cosmin = min(max(float(v), 0.0), 1.0)

    cosmin = property(lambda self: self._cosmin, _set_cosmin)

    def _set_direction(self, d):
        d = np.array(d, dtype=np.float64)
        norm = np.linalg.norm(d)
        if norm == 0.0:
            raise ValueError('Direction vector norm/length must not be 0!')
        self._direction = d*(1.0/norm)

    direction = property(lambda self: self._direction, _set_direction)


def _inv_step(axis):
    return 1.0/axis.step if axis.step != 0.0 else 0.0


class Total(Detector):
    cu_type = 'xo::DetTotal'

    def cl_type(self, mc):
        T = mc.types
        class ClTotal(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('cos_min', T.mc_fp_t),
                        ('offset', T.mc_size_t)]
        return ClTotal

    def __init__(self, cosmin=0.0, direction=(0.0, 0.0, 1.0)):
        if isinstance(cosmin, Total):
            o = cosmin
            cosmin, direction = o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((1,)), 0
        super().__init__(raw, nphotons)
        self.cosmin, self.direction = cosmin, direction

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.cos_min = self._cosmin
        target.direction.fromarray(self._direction)
        return target

    def todict(self):
        return {'type': 'Total', 'cosmin': self._cosmin,
                'direction': self._direction.tolist()}


class TotalLut(Detector):
    """Total reflectance / transmittance weighted by an angular sensitivity table
    (total.py:240-460)."""
    cu_type = 'xo::DetTotalLut'

    def cl_type(self, mc):
        T = mc.types
        class ClTotalLut(cltypes.Structure):
            _fields_ = [('lut', CollectionLut.cl_type(mc)), ('direction', T.mc_point3f_t),
                        ('offset', T.mc_size_t)]
        return ClTotalLut

    def __init__(self, lut, direction=(0.0, 0.0, 1.0)):
        if isinstance(lut, TotalLut):
            o = lut
            lut, direction = o.lut, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((1,)), 0
        super().__init__(raw, nphotons)
        self.lut = lut
        self.direction = direction

    def _set_lut(self, lut):
        if isinstance(lut, LinearLut) and not isinstance(lut, CollectionLut):
            lut = CollectionLut(lut)        # (a deserialised table: lut.py:209-214 drops the subclass)
        if not isinstance(lut, CollectionLut):
            raise TypeError('The lookup table must be an instance of CollectionLut!')
        self._lut = lut

    lut = property(lambda self: self._lut, _set_lut)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.lut = self._lut.cl_pack(mc, target.lut)
        target.direction.fromarray(self._direction)
        return target

    def todict(self):
        return {'type': 'TotalLut', 'lut': self._lut.todict(),
                'direction': self._direction.tolist()}


class TotalLutPl(TotalLut):
    """TotalLut resolved by optical path length (totalpl.py:315-611)."""
    cu_type = 'xo::DetTotalLutPl'

    def cl_type(self, mc):
        T = mc.types
        class ClTotalLutPl(cltypes.Structure):
            _fields_ = [('lut', CollectionLut.cl_type(mc)), ('direction', T.mc_point3f_t),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('pl_log_scale', T.mc_int_t)]
        return ClTotalLutPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, lut, plaxis=None, direction=(0.0, 0.0, 1.0)):
        if isinstance(lut, TotalLutPl):
            o = lut
            super().__init__(o.lut, o.direction)
            plaxis = type(o.plaxis)(o.plaxis)
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            super().__init__(lut, direction)
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n,)), 0
        self._raw_data, self._nphotons = raw, int(nphotons)
        self._pl_axis = plaxis

    plaxis = property(lambda self: self._pl_axis)
    pl = property(lambda self: self._pl_axis.centers)
    pledges = property(lambda self: self._pl_axis.edges)
    npl = property(lambda self: self._pl_axis.n)

    def cl_pack(self, mc, target=None):
        target = super().cl_pack(mc, target)
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        return {'type': 'TotalLutPl', 'lut': self._lut.todict(),
                'plaxis': self._pl_axis.todict(), 'direction': self._direction.tolist()}


class Radial(Detector):
    cu_type = 'xo::DetRadial'

    def cl_type(self, mc):
        T = mc.types
        class ClRadial(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('position', T.mc_point2f_t),
                        ('r_min', T.mc_fp_t), ('inv_dr', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n', T.mc_size_t),
                        ('offset', T.mc_size_t), ('log_scale', T.mc_int_t)]
        return ClRadial

    def __init__(self, raxis, position=(0.0, 0.0), cosmin=0.0,
                 direction=(0.0, 0.0, 1.0)):
        if isinstance(raxis, Radial):
            o = raxis
            position, cosmin, direction = o.position, o.cosmin, o.direction
            raxis = type(o.raxis)(o.raxis)
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((raxis.n,)), 0
        super().__init__(raw, nphotons)
        self._r_axis = raxis
        self._position = np.zeros((2,))
        self._position[:] = position
        self.cosmin, self.direction = cosmin, direction
        e = raxis.edges
        self._inv_accumulators_area = 1.0/(np.pi*(e[1:]**2 - e[:-1]**2))

    raxis = property(lambda self: self._r_axis)
    r = property(lambda self: self._r_axis.centers)
    edges = property(lambda self: self._r_axis.edges)
    n = property(lambda self: self._r_axis.n)
    logscale = property(lambda self: self._r_axis.logscale)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    @property
    def normalized(self):
        return self.raw*self._inv_accumulators_area*(1.0/max(self._nphotons, 1))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.position.fromarray(self._position)
        target.r_min = self._r_axis.scaled_start
        target.inv_dr = _inv_step(self._r_axis)
        target.log_scale = self._r_axis.logscale
        target.n = self._r_axis.n
        target.cos_min = self._cosmin
        target.direction.fromarray(self._direction)
        return target

    def todict(self):
        return {'type': 'Radial', 'position': self._position.tolist(),
                'r_axis': self._r_axis.todict(), 'cosmin': self._cosmin,
                'direction': self._direction.tolist()}


class Cartesian(Detector):
    """x-y grid; raw data indexed [y, x] (cartesian.py)."""
    cu_type = 'xo::DetCartesian'

    def cl_type(self, mc):
        T = mc.types
        class ClCartesian(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('x_min', T.mc_fp_t),
                        ('inv_dx', T.mc_fp_t), ('y_min', T.mc_fp_t),
                        ('inv_dy', T.mc_fp_t), ('cos_min', T.mc_fp_t),
                        ('n_x', T.mc_size_t), ('n_y', T.mc_size_t),
                        ('offset', T.mc_size_t)]
        return ClCartesian

    def __init__(self, xaxis, yaxis=None, cosmin=0.0, direction=(0.0, 0.0, 1.0)):
        if isinstance(xaxis, Cartesian):
            o = xaxis
            xaxis, yaxis = type(o.xaxis)(o.xaxis), type(o.yaxis)(o.yaxis)
            cosmin, direction = o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if yaxis is None:
                yaxis = Axis(xaxis)
            raw, nphotons = np.zeros((yaxis.n, xaxis.n)), 0
        super().__init__(raw, nphotons)
        self._x_axis, self._y_axis = xaxis, yaxis
        self.cosmin, self.direction = cosmin, direction
        self._accumulators_area = abs(xaxis.step*yaxis.step)

    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)

    @property
    def normalized(self):
        return self.raw*(1.0/(max(self.nphotons, 1.0)*self._accumulators_area))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.direction.fromarray(self._direction)
        target.x_min, target.inv_dx = self._x_axis.start, _inv_step(self._x_axis)
        target.y_min, target.inv_dy = self._y_axis.start, _inv_step(self._y_axis)
        target.cos_min = self._cosmin
        target.n_x, target.n_y = self._x_axis.n, self._y_axis.n
        return target

    def todict(self):
        return {'type': 'Cartesian', 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'cosmin': self._cosmin,
                'direction': self._direction.tolist()}


class SixAroundOne(Detector):
    """Six-around-one fiber probe; raw[0] central fiber, raw[1..6] ring
    counter-clockwise from +x (probe/sixaroundone.py)."""
    cu_type = 'xo::DetSixAroundOne'

    def cl_type(self, mc):
        T = mc.types
        class ClSixAroundOne(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point2f_t),
                        ('core_r_squared', T.mc_fp_t), ('core_spacing', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('offset', T.mc_size_t)]
        return ClSixAroundOne

    def __init__(self, fiber, spacing: float = None, position=(0.0, 0.0),
                 direction=(0.0, 0.0, 1.0)):
        if isinstance(fiber, SixAroundOne):
            o = fiber
            fiber, spacing, position, direction = o.fiber, o.spacing, o.position, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((7,)), 0
            if spacing is None:
                spacing = fiber.dcladding
        super().__init__(raw, nphotons)
        self._fiber = fiber
        self._spacing = float(spacing)
        self._position = np.zeros((2,))
        self._position[:] = position
        self.direction = direction

    fiber = property(lambda self: self._fiber)
    spacing = property(lambda self: self._spacing)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def fiber_position(self, index: int) -> Tuple[float, float]:
        if index >= 7 or index < -7:
            raise IndexError('The fiber index is out of valid range!')
        index %= 7
        if index == 0:
            return (0.0, 0.0)
        ang = (index - 1)*np.pi/3.0
        return (self._spacing*np.cos(ang), self._spacing*np.sin(ang))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        adir = self._direction[0], self._direction[1], abs(self._direction[2])
        target.transformation.fromarray(geometry.transform_base(adir, (0.0, 0.0, 1.0)))
        target.core_spacing = self._spacing
        target.core_r_squared = 0.25*self._fiber.dcore**2
        target.cos_min = (1.0 - (self._fiber.na/self._fiber.ncore)**2)**0.5
        target.position.fromarray(self._position)
        return target

    def todict(self):
        return {'type': 'SixAroundOne', 'fiber': self._fiber.todict(),
                'spacing': self._spacing, 'position': self._position.tolist(),
                'direction': self._direction.tolist()}


class LinearArray(Detector):
    """Linear array of ``n`` equally spaced fibers (probe/lineararray.py); raw[i] is
    fiber i, counted along ``orientation``.  ``n`` is a compile-time feature."""
    def cu_type(self, mc):
        return 'xo::DetLinearArray<{}>'.format(self._n)

    def cl_type(self, mc):
        T = mc.types
        class ClLinearArray(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t),
                        ('first_position', T.mc_point2f_t),
                        ('delta_position', T.mc_point2f_t),
                        ('core_r_squared', T.mc_fp_t), ('cos_min', T.mc_fp_t),
                        ('offset', T.mc_size_t)]
        return ClLinearArray

    def __init__(self, fiber, n: int = 1, spacing: float = None, orientation=(1.0, 0.0),
                 position=(0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        if isinstance(fiber, LinearArray):
            o = fiber
            fiber, n, spacing = o.fiber, o.n, o.spacing
            orientation, position, direction = o.orientation, o.position, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            n = max(int(n), 1)
            raw, nphotons = np.zeros((n,)), 0
            if spacing is None:
                spacing = fiber.dcladding
        super().__init__(raw, nphotons)
        self._fiber = fiber
        self._n = max(int(n), 1)
        self._spacing = float(spacing)
        self._orientation = np.array((1.0, 0.0))
        self._position = np.zeros((2,))
        self.orientation = orientation
        self.position = position
        self.direction = direction

    def _set_fiber(self, fiber):
        self._fiber = fiber

    fiber = property(lambda self: self._fiber, _set_fiber)
    n = property(lambda self: self._n)

    def _set_spacing(self, v):
        self._spacing = float(v)

    spacing = property(lambda self: self._spacing, _set_spacing)

    def _set_orientation(self, o):
        self._orientation[:] = o
        norm = np.linalg.norm(self._orientation)
        if norm == 0.0:
            raise ValueError('Orientation vector norm/length must not be 0!')
        self._orientation *= 1.0/norm

    orientation = property(lambda self: self._orientation, _set_orientation)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def check(self):
        if self._spacing < self._fiber.dcore:
            raise ValueError('Spacing between the optical fibers is smaller '
                             'than the diameter of the fiber core!')
        return True

    def fiber_position(self, index: int) -> Tuple[float, float]:
        if index >= self._n or index < -self._n:
            raise IndexError('The fiber index is out of valid range!')
        left = self._position - self._orientation*self._spacing*(self._n - 1)*0.5
        return tuple(left + self._spacing*self._orientation*int(index))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        adir = self._direction[0], self._direction[1], abs(self._direction[2])
        target.transformation.fromarray(geometry.transform_base(adir, (0.0, 0.0, 1.0)))
        target.core_r_squared = 0.25*self._fiber.dcore**2
        target.cos_min = (1.0 - (self._fiber.na/self._fiber.ncore)**2)**0.5
        target.first_position.fromarray(self.fiber_position(0))
        target.delta_position.fromarray(self._orientation*self._spacing)
        return target

    def todict(self):
        return {'type': 'LinearArray', 'fiber': self._fiber.todict(), 'n': self._n,
                'orientation': self._orientation.tolist(), 'spacing': self._spacing,
                'position': self._position.tolist(), 'direction': self._direction.tolist()}

    @staticmethod
    def fromdict(data):
        data = dict(data)
        if data.pop('type') != 'LinearArray':
            raise TypeError('Expected a "LinearArray" type!')
        return LinearArray(MultimodeFiber.fromdict(data.pop('fiber')), **data)


class FiberArray(Detector):
    """Array of individually placed / tilted fibers (probe/fiberarray.py); raw[i] is
    fiber i of the list.  The number of fibers is a compile-time feature."""
    def cu_type(self, mc):
        return 'xo::DetFiberArray<{}>'.format(len(self._fibers))

    def cl_type(self, mc):
        T = mc.types
        n = self.n
        class ClFiberArray(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t*n),
                        ('core_position', T.mc_point2f_t*n),
                        ('core_r_squared', T.mc_fp_t*n), ('cos_min', T.mc_fp_t*n),
                        ('offset', T.mc_size_t)]
        return ClFiberArray

    def __init__(self, fibers):
        if isinstance(fibers, FiberArray):
            o = fibers
            fibers, raw, nphotons = o.fibers, np.copy(o.raw), o.nphotons
        else:
            fibers = list(fibers)
            raw, nphotons = np.zeros((len(fibers),)), 0
        super().__init__(raw, nphotons)
        self._fibers = fibers

    def _set_fibers(self, fibers):
        if len(self._fibers) != len(fibers):
            raise ValueError('The number of optical fibers must not change!')
        if type(self._fibers[0]) != type(fibers[0]):
            raise TypeError('The type of optical fibers must not change!')
        self._fibers[:] = fibers

    fibers = property(lambda self: self._fibers, _set_fibers)
    n = property(lambda self: len(self._fibers))

    def __len__(self):
        return len(self._fibers)

    def __iter__(self):
        return iter(self._fibers)

    def __getitem__(self, what):
        return self._fibers[what]

    def check(self):
        for a in self._fibers:
            for b in self._fibers:
                if a is not b:
                    d = np.linalg.norm(a.position - b.position)
                    if d < max(a.fiber.dcladding, b.fiber.dcladding):
                        raise ValueError('Some of the fibers in the detector array overlap!')
        return True

    def fiber_position(self, index: int) -> Tuple[float, float]:
        n = len(self._fibers)
        if index >= n or index < -n:
            raise IndexError('The fiber index is out of valid range!')
        return tuple(self._fibers[index].position[:2])

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        for index, cfg in enumerate(self._fibers):
            adir = cfg.direction[0], cfg.direction[1], abs(cfg.direction[2])
            target.transformation[index].fromarray(
                geometry.transform_base(adir, (0.0, 0.0, 1.0)))
            target.core_position[index].fromarray(cfg.position)
            target.core_r_squared[index] = 0.25*cfg.fiber.dcore**2
            target.cos_min[index] = (1.0 - (cfg.fiber.na/cfg.fiber.ncore)**2)**0.5
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        return target

    def todict(self):
        return {'type': 'FiberArray', 'fibers': [f.todict() for f in self._fibers]}

    @staticmethod
    def fromdict(data):
        data = dict(data)
        if data.pop('type') != 'FiberArray':
            raise TypeError('Expected a "FiberArray" type!')
        return FiberArray([FiberLayout.fromdict(f) for f in data.pop('fibers')])


class LinearArrayPl(LinearArray):
    """LinearArray resolved by optical path length; raw indexed [pl, fiber]
    (probe/lineararraypl.py)."""
    def cu_type(self, mc):
        return 'xo::DetLinearArrayPl<{}>'.format(self._n)

    def cl_type(self, mc):
        T = mc.types
        class ClLinearArrayPl(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t),
                        ('first_position', T.mc_point2f_t),
                        ('delta_position', T.mc_point2f_t),
                        ('core_r_squared', T.mc_fp_t), ('cos_min', T.mc_fp_t),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('pl_log_scale', T.mc_int_t)]
        return ClLinearArrayPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, fiber, n: int = 1, spacing: float = None, plaxis=None,
                 orientation=(1.0, 0.0), position=(0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        if isinstance(fiber, LinearArrayPl):
            o = fiber
            super().__init__(o.fiber, o.n, o.spacing, o.orientation, o.position, o.direction)
            plaxis = type(o.plaxis)(o.plaxis)
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            super().__init__(fiber, n, spacing, orientation, position, direction)
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n, self._n)), 0
        self._raw_data, self._nphotons = raw, int(nphotons)
        self._pl_axis = plaxis

    plaxis = property(lambda self: self._pl_axis)
    pl = property(lambda self: self._pl_axis.centers)
    pledges = property(lambda self: self._pl_axis.edges)
    npl = property(lambda self: self._pl_axis.n)

    def cl_pack(self, mc, target=None):
        target = super().cl_pack(mc, target)
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        d = super().todict()
        d.update(type='LinearArrayPl', pl_axis=self._pl_axis.todict())
        return d

    @staticmethod
    def fromdict(data):
        data = dict(data)
        if data.pop('type') != 'LinearArrayPl':
            raise TypeError('Expected a "LinearArrayPl" type!')
        pl = dict(data.pop('pl_axis'))
        pl.pop('type', None)
        return LinearArrayPl(MultimodeFiber.fromdict(data.pop('fiber')), plaxis=Axis(**pl), **data)


class FiberArrayPl(FiberArray):
    """FiberArray resolved by optical path length; raw indexed [pl, fiber]
    (probe/fiberarraypl.py)."""
    def cu_type(self, mc):
        return 'xo::DetFiberArrayPl<{}>'.format(len(self._fibers))

    def cl_type(self, mc):
        T = mc.types
        n = self.n
        class ClFiberArrayPl(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t*n),
                        ('core_position', T.mc_point2f_t*n),
                        ('core_r_squared', T.mc_fp_t*n), ('cos_min', T.mc_fp_t*n),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('pl_log_scale', T.mc_int_t)]
        return ClFiberArrayPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, fibers, plaxis=None):
        if isinstance(fibers, FiberArrayPl):
            o = fibers
            super().__init__(o.fibers)
            plaxis = type(o.plaxis)(o.plaxis)
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            super().__init__(fibers)
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n, len(self._fibers))), 0
        self._raw_data, self._nphotons = raw, int(nphotons)
        self._pl_axis = plaxis

    plaxis = property(lambda self: self._pl_axis)
    pl = property(lambda self: self._pl_axis.centers)
    pledges = property(lambda self: self._pl_axis.edges)
    npl = property(lambda self: self._pl_axis.n)

    def cl_pack(self, mc, target=None):
        target = super().cl_pack(mc, target)
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        return {'type': 'FiberArrayPl', 'fibers': [f.todict() for f in self._fibers],
                'pl_axis': self._pl_axis.todict()}


class FiberLutArray(FiberArray):
    """Array of fibers (``MultimodeFiberLut``) with tabulated collection
    sensitivity (probe/fiberlutarray.py)."""
    def cu_type(self, mc):
        return 'xo::DetFiberLutArray<{}>'.format(len(self._fibers))

    def cl_type(self, mc):
        T = mc.types
        n = self.n
        class ClFiberLutArray(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t*n),
                        ('core_position', T.mc_point2f_t*n),
                        ('core_r_squared', T.mc_fp_t*n),
                        ('lut', CollectionLut.cl_type(mc)*n),
                        ('offset', T.mc_size_t)]
        return ClFiberLutArray

    def cl_options(self, mc):
        return [('MC_USE_FP_LUT', True)]

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        for index, cfg in enumerate(self._fibers):
            adir = cfg.direction[0], cfg.direction[1], abs(cfg.direction[2])
            target.transformation[index].fromarray(
                geometry.transform_base(adir, (0.0, 0.0, 1.0)))
            target.core_position[index].fromarray(cfg.position)
            target.core_r_squared[index] = 0.25*cfg.fiber.dcore**2
            cfg.fiber.collection.cl_pack(mc, target.lut[index])
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        return target

    def todict(self):
        return {'type': 'FiberLutArray', 'fibers': [f.todict() for f in self._fibers]}


class RadialPl(Detector):
    """Radial x optical-path-length histogram; raw indexed [pl, r] (radialpl.py)."""
    cu_type = 'xo::DetRadialPl'

    def cl_type(self, mc):
        T = mc.types
        class ClRadialPl(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('position', T.mc_point2f_t),
                        ('r_min', T.mc_fp_t), ('inv_dr', T.mc_fp_t),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n_r', T.mc_size_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('r_log_scale', T.mc_int_t), ('pl_log_scale', T.mc_int_t)]
        return ClRadialPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, raxis, plaxis=None, position=(0.0, 0.0), cosmin=0.0,
                 direction=(0.0, 0.0, 1.0)):
        if isinstance(raxis, RadialPl):
            o = raxis
            raxis, plaxis = type(o.raxis)(o.raxis), type(o.plaxis)(o.plaxis)
            position, cosmin, direction = o.position, o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n, raxis.n)), 0
        super().__init__(raw, nphotons)
        self._r_axis, self._pl_axis = raxis, plaxis
        self._position = np.zeros((2,))
        self._position[:] = position
        self.cosmin, self.direction = cosmin, direction
        e = raxis.edges
        self._inv_accumulators_area = 1.0/(np.pi*(e[1:]**2 - e[:-1]**2))

    raxis = property(lambda self: self._r_axis)
    plaxis = property(lambda self: self._pl_axis)
    r = property(lambda self: self._r_axis.centers)
    pl = property(lambda self: self._pl_axis.centers)
    position = property(lambda self: self._position)

    @property
    def normalized(self):
        return self.raw*self._inv_accumulators_area*(1.0/max(self.nphotons, 1.0))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.position.fromarray(self._position)
        target.r_min, target.inv_dr = self._r_axis.scaled_start, _inv_step(self._r_axis)
        target.r_log_scale, target.n_r = self._r_axis.logscale, self._r_axis.n
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        target.cos_min = self._cosmin
        target.direction.fromarray(self._direction)
        return target

    def todict(self):
        return {'type': 'RadialPl', 'position': self._position.tolist(),
                'raxis': self._r_axis.todict(), 'plaxis': self._pl_axis.todict(),
                'cosmin': self._cosmin, 'direction': self._direction.tolist()}


class TotalPl(Detector):
    """Optical-path-length histogram of all escaping packets (totalpl.py)."""
    cu_type = 'xo::DetTotalPl'

    def cl_type(self, mc):
        T = mc.types
        class ClTotalPl(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('cos_min', T.mc_fp_t),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('pl_log_scale', T.mc_int_t)]
        return ClTotalPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, plaxis=None, cosmin=0.0, direction=(0.0, 0.0, 1.0)):
        if isinstance(plaxis, TotalPl):
            o = plaxis
            plaxis, cosmin, direction = type(o.plaxis)(o.plaxis), o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n,)), 0
        super().__init__(raw, nphotons)
        self._pl_axis = plaxis
        self.cosmin, self.direction = cosmin, direction

    plaxis = property(lambda self: self._pl_axis)
    pl = property(lambda self: self._pl_axis.centers)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.cos_min = self._cosmin
        target.direction.fromarray(self._direction)
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        return {'type': 'TotalPl', 'plaxis': self._pl_axis.todict(),
                'cosmin': self._cosmin, 'direction': self._direction.tolist()}


class SymmetricX(Detector):
    """Bins along x, symmetric around a center; integrates over y (symmetric.py)."""
    cu_type = 'xo::DetSymmetricX'

    def cl_type(self, mc):
        T = mc.types
        class ClSymmetricX(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('position_x', T.mc_fp_t),
                        ('x_offset', T.mc_fp_t), ('inv_step', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n_half', T.mc_size_t),
                        ('log_scale', T.mc_int_t), ('offset', T.mc_size_t)]
        return ClSymmetricX

    def __init__(self, xaxis, cosmin: float = 0.0, direction=(0.0, 0.0, 1.0)):
        if isinstance(xaxis, SymmetricX):
            o = xaxis
            xaxis, cosmin, direction = type(o.xaxis)(o.xaxis), o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            raw, nphotons = np.zeros((xaxis.n,)), 0
        super().__init__(raw, nphotons)
        self._x_axis = xaxis
        self.cosmin, self.direction = cosmin, direction
        self._inv_accumulators_width = 1.0/(xaxis.edges[1:] - xaxis.edges[:-1])

    xaxis = property(lambda self: self._x_axis)
    x = property(lambda self: self._x_axis.centers)
    edges = property(lambda self: self._x_axis.edges)

    @property
    def normalized(self):
        return self.raw*self._inv_accumulators_width*(1.0/max(self.nphotons, 1.0))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.direction.fromarray(self._direction)
        target.position_x = self._x_axis.center
        target.x_offset = self._x_axis.scaled_offset
        target.inv_step = 1.0/self._x_axis.step if self._x_axis.step != 0.0 else 0.0
        target.log_scale = self._x_axis.logscale
        target.n_half = self._x_axis.n_half
        target.cos_min = self._cosmin
        return target

    def todict(self):
        return {'type': 'SymmetricX', 'xaxis': self._x_axis.todict(), 'cosmin': self._cosmin,
                'direction': self._direction.tolist()}


class CartesianPl(Detector):
    """x-y grid x optical-path-length histogram; raw indexed [pl, y, x]
    (cartesianpl.py)."""
    cu_type = 'xo::DetCartesianPl'

    def cl_type(self, mc):
        T = mc.types
        class ClCartesianPl(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('x_min', T.mc_fp_t),
                        ('inv_dx', T.mc_fp_t), ('y_min', T.mc_fp_t),
                        ('inv_dy', T.mc_fp_t), ('pl_min', T.mc_fp_t),
                        ('inv_dpl', T.mc_fp_t), ('cos_min', T.mc_fp_t),
                        ('n_x', T.mc_size_t), ('n_y', T.mc_size_t),
                        ('n_pl', T.mc_size_t), ('offset', T.mc_size_t),
                        ('pl_log_scale', T.mc_int_t)]
        return ClCartesianPl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, xaxis, yaxis=None, plaxis=None, cosmin=0.0,
                 direction=(0.0, 0.0, 1.0)):
        if isinstance(xaxis, CartesianPl):
            o = xaxis
            xaxis, yaxis = type(o.xaxis)(o.xaxis), type(o.yaxis)(o.yaxis)
            plaxis = type(o.plaxis)(o.plaxis)
            cosmin, direction = o.cosmin, o.direction
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if yaxis is None:
                yaxis = Axis(xaxis)
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            raw, nphotons = np.zeros((plaxis.n, yaxis.n, xaxis.n)), 0
        super().__init__(raw, nphotons)
        self._x_axis, self._y_axis, self._pl_axis = xaxis, yaxis, plaxis
        self.cosmin, self.direction = cosmin, direction
        self._accumulators_area = abs(xaxis.step*yaxis.step)

    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    plaxis = property(lambda self: self._pl_axis)
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)
    pl = property(lambda self: self._pl_axis.centers)

    @property
    def normalized(self):
        return self.raw*(1.0/(max(self.nphotons, 1.0)*self._accumulators_area))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.direction.fromarray(self._direction)
        target.x_min = self._x_axis.start
        target.inv_dx = 1.0/self._x_axis.step if self._x_axis.n > 1 else 0.0
        target.y_min = self._y_axis.start
        target.inv_dy = 1.0/self._y_axis.step if self._y_axis.n > 1 else 0.0
        target.cos_min = self._cosmin
        target.n_x, target.n_y = self._x_axis.n, self._y_axis.n
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        return {'type': 'CartesianPl', 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'plaxis': self._pl_axis.todict(),
                'cosmin': self._cosmin, 'direction': self._direction.tolist()}


class SixAroundOnePl(Detector):
    """Six-around-one fiber probe x optical-path-length histogram; raw indexed
    [pl, fiber] (probe/sixaroundonepl.py)."""
    cu_type = 'xo::DetSixAroundOnePl'

    def cl_type(self, mc):
        T = mc.types
        class ClSixAroundOnePl(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point2f_t),
                        ('core_r_squared', T.mc_fp_t), ('core_spacing', T.mc_fp_t),
                        ('pl_min', T.mc_fp_t), ('inv_dpl', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n_pl', T.mc_size_t),
                        ('offset', T.mc_size_t), ('pl_log_scale', T.mc_int_t)]
        return ClSixAroundOnePl

    def cl_options(self, mc):
        return [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    def __init__(self, fiber, spacing: float = None, plaxis=None, position=(0.0, 0.0),
                 direction=(0.0, 0.0, 1.0)):
        if isinstance(fiber, SixAroundOnePl):
            o = fiber
            fiber, spacing, position, direction = o.fiber, o.spacing, o.position, o.direction
            plaxis = type(o.plaxis)(o.plaxis)
            raw, nphotons = np.copy(o.raw), o.nphotons
        else:
            if plaxis is None:
                plaxis = Axis(0.0, 1.0, 1)
            if spacing is None:
                spacing = fiber.dcladding
            raw, nphotons = np.zeros((plaxis.n, 7)), 0
        super().__init__(raw, nphotons)
        self._fiber = fiber
        self._spacing = float(spacing)
        self._pl_axis = plaxis
        self._position = np.zeros((2,))
        self._position[:] = position
        self.direction = direction

    fiber = property(lambda self: self._fiber)
    spacing = property(lambda self: self._spacing)
    plaxis = property(lambda self: self._pl_axis)
    pl = property(lambda self: self._pl_axis.centers)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        adir = self._direction[0], self._direction[1], abs(self._direction[2])
        target.transformation.fromarray(geometry.transform_base(adir, (0.0, 0.0, 1.0)))
        target.core_spacing = self._spacing
        target.core_r_squared = 0.25*self._fiber.dcore**2
        target.cos_min = (1.0 - (self._fiber.na/self._fiber.ncore)**2)**0.5
        target.position.fromarray(self._position)
        target.pl_min, target.inv_dpl = self._pl_axis.scaled_start, _inv_step(self._pl_axis)
        target.pl_log_scale, target.n_pl = self._pl_axis.logscale, self._pl_axis.n
        return target

    def todict(self):
        return {'type': 'SixAroundOnePl', 'fiber': self._fiber.todict(),
                'spacing': self._spacing, 'plaxis': self._pl_axis.todict(),
                'position': self._position.tolist(), 'direction': self._direction.tolist()}


class Detectors(McObject):
    """Container {top, bottom, specular} (mcdetector/base.py:295-549)."""

    def __init__(self, top=None, bottom=None, specular=None):
        super().__init__()
        if isinstance(top, Detectors):
            d = top
            top, bottom, specular = (type(d.top)(d.top), type(d.bottom)(d.bottom),
                                     type(d.specular)(d.specular))
        top = DetectorDefault() if top is None else top
        bottom = DetectorDefault() if bottom is None else bottom
        specular = DetectorDefault() if specular is None else specular
        if isinstance(top, DetectorDefault) and top.location != NONE:
            top = DetectorDefault()
        if isinstance(bottom, DetectorDefault) and bottom.location != NONE:
            bottom = DetectorDefault()
        if isinstance(specular, DetectorDefault) and specular.location != NONE:
            specular = DetectorDefault()
        top.location, bottom.location, specular.location = TOP, BOTTOM, SPECULAR
        self._top, self._bottom, self._specular = top, bottom, specular

    top = property(lambda self: self._top)
    bottom = property(lambda self: self._bottom)
    specular = property(lambda self: self._specular)

    def cl_type(self, mc):
        class ClDetectors(cltypes.Structure):
            _fields_ = [('top', self._top.fetch_cl_type(mc)),
                        ('bottom', self._bottom.fetch_cl_type(mc)),
                        ('specular', self._specular.fetch_cl_type(mc))]
        return ClDetectors

    def cl_options(self, mc):
        options, used = [], False
        for det, name in ((self._top, 'TOP'), (self._bottom, 'BOTTOM'),
                          (self._specular, 'SPECULAR')):
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
        self._top.cl_pack(mc, target.top)
        self._bottom.cl_pack(mc, target.bottom)
        self._specular.cl_pack(mc, target.specular)
        return target

    def update_data(self, mc, detector, data, nphotons=0):
        location = detector.location if isinstance(detector, DetectorBase) else detector
        if location not in (TOP, BOTTOM, SPECULAR):
            raise ValueError('Detector location must be one of "{}", "{}" or "{}" '
                             'but got "{}"!'.format(TOP, BOTTOM, SPECULAR, location))
        getattr(self, location).update_data(
            mc, accumulators=data.get(np.dtype(mc.types.np_accu)),
            float_buffers=data.get(np.dtype(mc.types.np_float)),
            integer_buffers=data.get(np.dtype(mc.types.np_int)), nphotons=nphotons)

    def types(self):
        return type(self._top), type(self._bottom), type(self._specular)

    def __iter__(self):
        return iter([self._top, self._bottom, self._specular])

    def todict(self):
        return {'type': 'Detectors', 'top': self._top.todict(),
                'bottom': self._bottom.todict(), 'specular': self._specular.todict()}
