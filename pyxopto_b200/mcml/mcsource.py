"""Photon packet sources of the planar simulator (mirror of ``xopto/mcml/mcsource``:
Line, GaussianBeam, UniformBeam, the fiber and rectangular sources, IsotropicPoint)."""
from typing import Tuple

import numpy as np

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil import boundary, geometry
from ..mcbase.mcutil.fiber import MultimodeFiber, MultimodeFiberLut  # noqa: F401
from ..mcbase.mcutil.lut import EmissionLut, LinearLut  # noqa: F401


class Source(McObject):
    def update(self, other):
        for key in self._update_keys:
            if isinstance(other, dict):
                if key in other:
                    setattr(self, key, other[key])
            else:
                setattr(self, key, getattr(other, key))

    _update_keys = ()


def _unit(v, name='direction'):
    v = np.array(v, dtype=np.float64)
    norm = np.linalg.norm(v)
    if norm == 0.0:
        raise ValueError('The norm/length of the propagation {} '
                         'vector must not be 0!'.format(name))
    return v/norm


class Line(Source):
    """Infinitely thin beam (mcsource/line.py)."""
    cu_type = 'xo::SrcLine'
    _update_keys = ('position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClLine(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('direction_medium', T.mc_point3f_t),
                        ('direction_sample', T.mc_point3f_t),
                        ('direction_reflected', T.mc_point3f_t),
                        ('reflectance', T.mc_fp_t)]
        return ClLine

    def __init__(self, position=(0.0, 0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.zeros((3,))
        self.position = position
        self.direction = direction

    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        self._direction[:] = _unit(d)

    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        d = self._direction
        t = -self._position[2]/d[2]
        position = self._position + d*t
        position[2] = 0.0
        n1, n2 = mc.layer(0).n, mc.layer(1).n
        reflectance = boundary.reflectance(n1, n2, d[2])
        if reflectance >= 1.0:
            raise ValueError('The line source is fully reflected from the top '
                             'surface of the sample!')
        sin1 = (1.0 - d[2]**2)**0.5
        sin2 = max(min(n1*sin1/n2, 1.0), -1.0)
        refracted = (d[0]*n1/n2, d[1]*n1/n2, np.sign(d[2])*(1 - sin2**2)**0.5)
        target.position.fromarray(position)
        target.direction_medium.fromarray(d)
        target.direction_sample.fromarray(refracted)
        target.direction_reflected.fromarray((d[0], d[1], -d[2]))
        target.reflectance = reflectance
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class GaussianBeam(Source):
    """Collimated Gaussian beam (mcsource/gaussianbeam.py)."""
    cu_type = 'xo::SrcGaussianBeam'
    cu_refill_lanes = 4     # long launch path: launch jointly (mcsim._refill_lanes)
    _update_keys = ('sigma', 'clip', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClGaussianBeam(cltypes.Structure):
            _pack_ = 1
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('sigma', T.mc_point2f_t),
                        ('clip', T.mc_fp_t), ('reflectance', T.mc_fp_t)]
        return ClGaussianBeam

    def __init__(self, sigma, clip: float = 5.0, position=(0.0, 0.0, 0.0),
                 direction=(0.0, 0.0, 1.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self._sigma = np.zeros((2,))
        self.sigma, self.clip = sigma, clip
        self.position, self.direction = position, direction

    def _set_sigma(self, s):
        self._sigma[:] = s
        if np.any(self._sigma < 0.0):
            raise ValueError('Beam diameter/sigma must not be negative!')

    def _set_clip(self, c):
        self._clip = float(c)

    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        d = _unit(d)
        if d[-1] <= 0.0:
            raise ValueError('Z component of the propagation direction '
                             'must be positive!')
        self._direction[:] = d

    sigma = property(lambda self: self._sigma, _set_sigma)
    clip = property(lambda self: self._clip, _set_clip)
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    @staticmethod
    def fwhm2sigma(fwhm: float) -> float:
        return fwhm/(8*np.log(2))**0.5

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        k = (0.0 - self._position[2])/self._direction[2]
        position = self._position + k*self._direction
        position[2] = 0.0
        n1, n2 = mc.layers[0].n, mc.layers[1].n
        direction = boundary.refract(self._direction, (0.0, 0.0, 1.0), n1, n2)
        reflectance = boundary.reflectance(n1, n2, abs(self._direction[-1]))
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.position.fromarray(position)
        target.direction.fromarray(direction)
        target.sigma.fromarray(self._sigma)
        target.clip = self._clip
        target.reflectance = reflectance
        return target, None, None

    def todict(self):
        return {'sigma': self._sigma.tolist(), 'clip': self._clip,
                'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class UniformBeam(Source):
    """Collimated uniform (top-hat) beam with an elliptical cross section
    (mcsource/uniformbeam.py)."""
    cu_type = 'xo::SrcUniformBeam'
    cu_refill_lanes = 4
    _update_keys = ('diameter', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformBeam(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_point2f_t),
                        ('reflectance', T.mc_fp_t)]
        return ClUniformBeam

    def __init__(self, diameter, position=(0.0, 0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self._diameter = np.zeros((2,))
        self.diameter = diameter
        self.position, self.direction = position, direction

    def _set_diameter(self, d):
        self._diameter[:] = d
        self._diameter = np.maximum(0.0, self._diameter)

    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        d = _unit(d)
        if d[-1] <= 0.0:
            raise ValueError('Z component of the propagation direction '
                             'must be positive!')
        self._direction[:] = d

    diameter = property(lambda self: self._diameter, _set_diameter, None,
                        'Beam diameter along the x and y axis (m).')
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        k = (0.0 - self._position[2])/self._direction[2]
        position = self._position + k*self._direction
        position[2] = 0.0
        n1, n2 = mc.layers[0].n, mc.layers[1].n
        direction = boundary.refract(self._direction, (0.0, 0.0, 1.0), n1, n2)
        reflectance = boundary.reflectance(n1, n2, abs(self._direction[-1]))
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.position.fromarray(position)
        target.direction.fromarray(direction)
        target.radius.x = self._diameter[0]*0.5
        target.radius.y = self._diameter[1]*0.5
        target.reflectance = reflectance
        return target, None, None

    def todict(self):
        return {'diameter': self._diameter.tolist(), 'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class UniformFiber(Source):
    """Optical fiber with uniform emission within the NA (mcsource/fiber.py)."""
    cu_type = 'xo::SrcUniformFiber'
    cu_refill_lanes = 8     # waiting lanes that trigger a service round (mcsim._refill_lanes)
    _update_keys = ('fiber', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformFiber(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n', T.mc_fp_t)]
        return ClUniformFiber

    def __init__(self, fiber: MultimodeFiber, position=(0.0, 0.0, 0.0),
                 direction=(0.0, 0.0, 1.0)):
        super().__init__()
        self._fiber = fiber
        self._position = np.zeros((3,))
        self._direction = np.zeros((3,))
        self.position, self.direction = position, direction

    def _set_fiber(self, f):
        self._fiber = f

    def _set_position(self, p):
        self._position[:] = p
        self._position[2] = 0.0

    def _set_direction(self, d):
        norm = np.linalg.norm(d)
        if norm == 0.0:
            raise ValueError('The direction vector is singular!')
        d = np.asarray(d, dtype=np.float64)*(1.0/norm)
        if d[-1] <= 0.0:
            raise ValueError('Z component of the propagation direction '
                             'must be positive!')
        self._direction[:] = d

    fiber = property(lambda self: self._fiber, _set_fiber)
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.n = self._fiber.ncore
        target.cos_min = (1.0 - self._fiber.na**2)**0.5
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.radius = self._fiber.dcore*0.5
        return target, None, None

    def todict(self):
        return {'fiber': self._fiber.todict(), 'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class LambertianFiber(UniformFiber):
    """Optical fiber emitting a lambertian beam within the NA (mcsource/fiber.py:499-688)."""
    cu_type = 'xo::SrcLambertianFiber'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClLambertianFiber(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('na', T.mc_fp_t), ('n', T.mc_fp_t)]
        return ClLambertianFiber

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.n = self._fiber.ncore
        target.na = self._fiber.na
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.radius = self._fiber.dcore*0.5
        return target, None, None


class UniformFiberLut(UniformFiber):
    """Optical fiber with a tabulated angular emission characteristic
    (mcsource/fiber.py:690-1017): the emission cosine comes from an EmissionLut in
    the float pool; unlike UniformFiber the direction *is* refracted into the
    sample."""
    cu_type = 'xo::SrcUniformFiberLut'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformFiberLut(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('n', T.mc_fp_t), ('lut', LinearLut.cl_type(mc))]
        return ClUniformFiberLut

    @staticmethod
    def cl_options(mc):
        return [('MC_USE_FP_LUT', True)]

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.radius = self._fiber.dcore*0.5
        target.n = self._fiber.ncore
        self._fiber.emission.cl_pack(mc, target.lut)
        return target, None, None


class UniformFiberNI(Source):
    """Optical fiber at normal incidence with uniform emission within the NA
    (mcsource/fiberni.py:180-419): no transformation, emission angle adjusted to
    the refractive index of the first sample layer."""
    cu_type = 'xo::SrcUniformFiberNI'
    cu_refill_lanes = 8
    _update_keys = ('fiber', 'position')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformFiberNI(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n', T.mc_fp_t)]
        return ClUniformFiberNI

    def __init__(self, fiber: MultimodeFiber, position=(0.0, 0.0, 0.0)):
        super().__init__()
        self._fiber = fiber
        self._position = np.zeros((3,))
        self.position = position

    def _set_fiber(self, f):
        self._fiber = f

    def _set_position(self, p):
        self._position[:] = p
        self._position[2] = 0.0

    fiber = property(lambda self: self._fiber, _set_fiber)
    position = property(lambda self: self._position, _set_position)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.n = self._fiber.ncore
        target.cos_min = (1.0 - self._fiber.na**2)**0.5
        target.position.fromarray(self._position)
        target.radius = self._fiber.dcore*0.5
        return target, None, None

    def todict(self):
        return {'fiber': self._fiber.todict(), 'position': self._position.tolist(),
                'type': type(self).__name__}

    @classmethod
    def fromdict(cls, data):
        data = dict(data)
        data.pop('type', None)
        fiber = data.pop('fiber')
        if isinstance(fiber, dict):
            kind = MultimodeFiberLut if 'emission' in fiber or 'collection' in fiber \
                else MultimodeFiber
            fiber = kind.fromdict(fiber)
        return cls(fiber, **data)

    def __str__(self):
        return '{} # id 0x{:>08X}.'.format(self.__repr__(), id(self))

    def __repr__(self):
        return '{}(fiber={}, position=({}, {}, {}))'.format(
            type(self).__name__, self._fiber, *self._position)


class LambertianFiberNI(UniformFiberNI):
    """Optical fiber at normal incidence emitting a lambertian beam within the NA
    (mcsource/fiberni.py:422-578)."""
    cu_type = 'xo::SrcLambertianFiberNI'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClLambertianFiberNI(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('na', T.mc_fp_t), ('n', T.mc_fp_t)]
        return ClLambertianFiberNI

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.n = self._fiber.ncore
        target.na = self._fiber.na
        target.position.fromarray(self._position)
        target.radius = self._fiber.dcore*0.5
        return target, None, None


class UniformFiberLutNI(UniformFiberNI):
    """Optical fiber at normal incidence with a tabulated angular emission
    characteristic (mcsource/fiberni.py:581-836; EmissionLut in the float pool)."""
    cu_type = 'xo::SrcUniformFiberLutNI'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformFiberLutNI(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('n', T.mc_fp_t), ('lut', LinearLut.cl_type(mc))]
        return ClUniformFiberLutNI

    @staticmethod
    def cl_options(mc):
        return [('MC_USE_FP_LUT', True)]

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.position.fromarray(self._position)
        target.radius = self._fiber.dcore*0.5
        target.n = self._fiber.ncore
        self._fiber.emission.cl_pack(mc, target.lut)
        return target, None, None


class UniformRectangular(Source):
    """Rectangular emitter (width x height) with uniform emission within the NA, at
    the top surface or inside a layer (mcsource/rectangular.py:32-315).  As in the
    reference the kernel cannot be built together with a specular detector."""
    cu_type = 'xo::SrcUniformRectangular'
    cu_refill_lanes = 4
    _update_keys = ('width', 'height', 'n', 'na', 'position')
    _aperture_field = 'cos_min'

    @classmethod
    def cl_type(cls, mc):
        T = mc.types
        class ClRectangular(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('size', T.mc_point2f_t),
                        ('n', T.mc_fp_t), ('cos_critical', T.mc_fp_t),
                        (cls._aperture_field, T.mc_fp_t), ('layer_index', T.mc_size_t)]
        return ClRectangular

    def __init__(self, width: float, height: float, n: float, na: float,
                 position=(0.0, 0.0, 0.0)):
        super().__init__()
        self._width, self._height = float(width), float(height)
        self._n, self._na = float(n), float(na)
        self._position = np.zeros((3,))
        self.position = position

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)
    width = property(lambda self: self._width, lambda self, v: setattr(self, '_width', float(v)))
    height = property(lambda self: self._height, lambda self, v: setattr(self, '_height', float(v)))
    n = property(lambda self: self._n, lambda self, v: setattr(self, '_n', float(v)))
    na = property(lambda self: self._na, lambda self, v: setattr(self, '_na', float(v)))

    def _aperture(self) -> float:
        return (1 - (self._na)**2)**0.5

    def cl_pack(self, mc, target=None):
        if mc.detectors is not None and type(mc.detectors.specular).__name__ != 'DetectorDefault':
            raise NotImplementedError(
                'Rectangular sources cannot be combined with a specular detector (the '
                'reference kernel does not build: rectangular.py:140 names a missing field).')
        if target is None:
            target = self.cl_type(mc)()
        if self._position[2] <= 0.0:
            position = (self._position[0], self._position[1], 0.0)
            layer_index = 1
        else:
            position = self._position
            layer_index = mc.layer_index(self._position[2])
        target.position.fromarray(position)
        target.size.fromarray([self._width, self._height])
        target.n = self._n
        target.cos_critical = boundary.cos_critical(self._n, mc.layers[layer_index].n)
        setattr(target, self._aperture_field, self._aperture())
        target.layer_index = layer_index
        return target, None, None

    def todict(self):
        return {'width': self._width, 'height': self._height, 'n': self._n, 'na': self._na,
                'position': self._position.tolist(), 'type': type(self).__name__}


class LambertianRectangular(UniformRectangular):
    """Rectangular emitter with lambertian emission within the NA
    (mcsource/rectangular.py:317-528)."""
    cu_type = 'xo::SrcLambertianRectangular'
    _aperture_field = 'na'

    def _aperture(self) -> float:
        return self._na


class UniformRectangularLut(Source):
    """Rectangular emitter with a tabulated angular emission characteristic
    (mcsource/rectangular.py:530-854): the emission cosine (valid for air) is
    sampled from an EmissionLut in the float pool and adjusted to the refractive
    index of the layer that holds the source."""
    cu_type = 'xo::SrcUniformRectangularLut'
    cu_refill_lanes = 4
    _update_keys = ('lut', 'width', 'height', 'n', 'position')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformRectangularLut(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('size', T.mc_point2f_t),
                        ('n', T.mc_fp_t), ('cos_critical', T.mc_fp_t),
                        ('lut', EmissionLut.cl_type(mc)), ('layer_index', T.mc_size_t)]
        return ClUniformRectangularLut

    @staticmethod
    def cl_options(mc):
        return [('MC_USE_FP_LUT', True)]

    def __init__(self, lut: EmissionLut, width: float, height: float, n: float,
                 position=(0.0, 0.0, 0.0)):
        super().__init__()
        self._lut = lut
        self._width, self._height = float(width), float(height)
        self._n = max(1.0, float(n))
        self._position = np.zeros((3,))
        self.position = position

    def _set_position(self, p):
        self._position[:] = p

    def _set_lut(self, lut):
        self._lut = lut

    position = property(lambda self: self._position, _set_position)
    lut = property(lambda self: self._lut, _set_lut)
    width = property(lambda self: self._width, lambda self, v: setattr(self, '_width', float(v)))
    height = property(lambda self: self._height, lambda self, v: setattr(self, '_height', float(v)))
    n = property(lambda self: self._n, lambda self, v: setattr(self, '_n', max(1.0, float(v))))

    def cl_pack(self, mc, target=None):
        if mc.detectors is not None and type(mc.detectors.specular).__name__ != 'DetectorDefault':
            raise NotImplementedError(
                'Rectangular sources cannot be combined with a specular detector (the '
                'reference kernel does not build: rectangular.py:650 names a missing field).')
        if target is None:
            target = self.cl_type(mc)()
        if self._position[2] <= 0.0:
            position = (self._position[0], self._position[1], 0.0)
            layer_index = 1
        else:
            position = self._position
            layer_index = mc.layer_index(self._position[2])
        self._lut.cl_pack(mc, target.lut)
        target.position.fromarray(position)
        target.size.fromarray([self._width, self._height])
        target.n = self._n
        target.cos_critical = boundary.cos_critical(self._n, mc.layers[layer_index].n)
        target.layer_index = layer_index
        return target, None, None

    def todict(self):
        return {'lut': self._lut.todict(), 'width': self._width, 'height': self._height,
                'n': self._n, 'position': self._position.tolist(),
                'type': type(self).__name__}


class IsotropicPoint(Source):
    """Isotropic point source above or inside the sample (mcsource/point.py)."""
    cu_type = 'xo::SrcIsotropicPoint'
    cu_refill_lanes = 4     # long launch path: launch jointly (mcsim._refill_lanes)
    _update_keys = ('position',)

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClIsotropicPoint(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('layer_index', T.mc_size_t)]
        return ClIsotropicPoint

    def __init__(self, position=(0.0, 0.0, 0.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self.position = position

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        if self._position[2] >= mc.layers.thickness():
            raise ValueError('The source must be located above or within '
                             'the sample but not under the sample!')
        target.position.fromarray(self._position)
        target.layer_index = 1 if self._position[2] <= 0.0 else \
            mc.layer_index(self._position[2])
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(), 'type': type(self).__name__}
