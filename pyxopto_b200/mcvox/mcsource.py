"""Photon packet sources of the voxelised simulator (mirror of
``xopto/mcvox/mcsource``: Line, GaussianBeam, UniformBeam, the fiber sources,
IsotropicPoint, IsotropicVoxel(s))."""
import numpy as np

from ..cl import cltypes
from ..mcbase.mcutil import boundary, geometry
from ..mcml.mcsource import Source, _unit
from ..mcbase.mcutil.fiber import MultimodeFiber, MultimodeFiberLut  # noqa: F401
from ..mcbase.mcutil.lut import EmissionLut, LinearLut  # noqa: F401


class Line(Source):
    """Infinitely thin beam entering the voxel box (mcvox/mcsource/line.py)."""
    cu_type = 'xo::VoxSrcLine'
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
        self.position, self.direction = position, direction

    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        self._direction[:] = _unit(d)

    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        search = -self._direction if mc.voxels.contains(self._position) else self._direction
        position, normal = mc.voxels.intersect(self._position, search)
        if position is None:
            raise ValueError('The line source does not intersect the voxelized sample!')
        medium, sample = mc.materials[0], mc.material(position)
        reflectance = boundary.reflectance(
            medium.n, sample.n, np.abs(np.dot(normal, self._direction)))
        if reflectance >= 1.0:
            raise ValueError('The line source is fully reflected from the '
                             'surface of the voxelized sample!')
        target.position.fromarray(position)
        target.direction_medium.fromarray(self._direction)
        target.direction_sample.fromarray(
            boundary.refract(self._direction, normal, medium.n, sample.n))
        target.direction_reflected.fromarray(boundary.reflect(self._direction, normal))
        target.reflectance = reflectance
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class GaussianBeam(Source):
    """Collimated Gaussian beam (mcvox/mcsource/gaussianbeam.py)."""
    cu_type = 'xo::VoxSrcGaussianBeam'
    cu_refill_lanes = 6     # long launch path: launch jointly (mcsim._refill_lanes)
    _update_keys = ('sigma', 'clip', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClGaussianBeam(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('sigma', T.mc_point2f_t),
                        ('clip', T.mc_fp_t)]
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
        self._direction[:] = _unit(d)

    sigma = property(lambda self: self._sigma, _set_sigma)
    clip = property(lambda self: self._clip, _set_clip)
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.sigma.fromarray(self._sigma)
        target.clip = self._clip
        return target, None, None

    def todict(self):
        return {'sigma': self._sigma.tolist(), 'clip': self._clip,
                'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class UniformBeam(Source):
    """Collimated uniform (top-hat) beam with an elliptical cross section
    (mcvox/mcsource/uniformbeam.py)."""
    cu_type = 'xo::VoxSrcUniformBeam'
    cu_refill_lanes = 6
    _update_keys = ('diameter', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformBeam(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_point2f_t)]
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
        self._direction[:] = _unit(d)

    diameter = property(lambda self: self._diameter, _set_diameter, None,
                        'Beam diameter along the x and y axis (m).')
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.radius.x = self._diameter[0]*0.5
        target.radius.y = self._diameter[1]*0.5
        return target, None, None

    def todict(self):
        return {'diameter': self._diameter.tolist(), 'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class UniformFiber(Source):
    """Multimode fibre with a uniform emission within its NA, pressed against
    (or pointing at) the voxel box (mcvox/mcsource/fiber.py UniformFiber)."""
    cu_type = 'xo::VoxSrcUniformFiber'
    cu_refill_lanes = 6
    _update_keys = ('fiber', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformFiber(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_fp_t),
                        ('cos_min', T.mc_fp_t), ('n', T.mc_fp_t)]
        return ClUniformFiber

    def __init__(self, fiber, position=(0.0, 0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        super().__init__()
        self._fiber = fiber
        self._position = np.zeros((3,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self.position, self.direction = position, direction

    def _set_fiber(self, f):
        self._fiber = f

    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        self._direction[:] = _unit(d)

    fiber = property(lambda self: self._fiber, _set_fiber)
    position = property(lambda self: self._position, _set_position)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.transformation.fromarray(
            geometry.transform_base((0.0, 0.0, 1.0), self._direction))
        target.n = self._fiber.ncore
        target.cos_min = (1.0 - (self._fiber.na)**2)**0.5
        target.position.fromarray(self._position)
        target.direction.fromarray(self._direction)
        target.radius = self._fiber.dcore*0.5
        return target, None, None

    def todict(self):
        return {'fiber': self._fiber.todict(), 'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': type(self).__name__}


class LambertianFiber(UniformFiber):
    """Multimode fibre emitting a lambertian beam within its NA
    (mcvox/mcsource/fiber.py:523-745)."""
    cu_type = 'xo::VoxSrcLambertianFiber'

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
    """Multimode fibre with a tabulated angular emission characteristic
    (mcvox/mcsource/fiber.py:747-): the emission cosine (valid for air) is sampled
    from the fibre's EmissionLut in the float pool."""
    cu_type = 'xo::VoxSrcUniformFiberLut'

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


class IsotropicPoint(Source):
    """Isotropic point source inside or outside the voxel box (mcvox/mcsource/point.py)."""
    cu_type = 'xo::VoxSrcIsotropicPoint'
    cu_refill_lanes = 6     # long launch path: launch jointly (mcsim._refill_lanes)
    _update_keys = ('position',)

    @staticmethod
    def cl_type(mc):
        class ClIsotropicPoint(cltypes.Structure):
            _fields_ = [('position', mc.types.mc_point3f_t)]
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
        target.position.fromarray(self._position)
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(), 'type': type(self).__name__}


class IsotropicVoxel(Source):
    """Isotropic emission from random points of one voxel (mcvox/mcsource/voxel.py:31-193)."""
    cu_type = 'xo::VoxSrcIsotropicVoxel'
    cu_refill_lanes = 6
    _update_keys = ('voxel',)

    @staticmethod
    def cl_type(mc):
        class ClIsotropicVoxel(cltypes.Structure):
            _fields_ = [('position', mc.types.mc_point3f_t), ('voxel', mc.types.mc_point3_t)]
        return ClIsotropicVoxel

    def __init__(self, voxel=(0, 0, 0)):
        super().__init__()
        self._voxel = np.zeros((3,), dtype=np.int32)
        self.voxel = voxel

    def _set_voxel(self, voxel):
        self._voxel[:] = voxel

    voxel = property(lambda self: self._voxel, _set_voxel, None,
                     'Source voxel indices (ind_x, ind_y, ind_z).')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        if not mc.voxels.isvalid(self._voxel):
            raise ValueError('Voxel index ({}, {}, {}) is not valid!'.format(*self._voxel))
        target.voxel.fromarray(self._voxel)
        target.position.fromarray(mc.voxels.center(self._voxel))
        return target, None, None

    def todict(self):
        return {'voxel': self._voxel.tolist(), 'type': type(self).__name__}


class IsotropicVoxels(Source):
    """Isotropic emission from a set of voxels with individual weights
    (mcvox/mcsource/voxel.py:195-420); the table (weight, x, y, z per voxel) lives in
    the simulator's float lookup-table pool."""
    cu_type = 'xo::VoxSrcIsotropicVoxels'
    cu_refill_lanes = 6

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClIsotropicVoxels(cltypes.Structure):
            _fields_ = [('position', T.mc_point3f_t), ('n', T.mc_size_t),
                        ('offset', T.mc_size_t)]
        return ClIsotropicVoxels

    @staticmethod
    def cl_options(mc):
        return [('MC_USE_FP_LUT', True)]

    def __init__(self, voxels, weights=None):
        super().__init__()
        voxels = np.asarray(voxels)
        if voxels.shape[-1] != 3:
            raise ValueError('Shape of the voxel array must be (n, 3)!')
        self._data = None
        self.voxels = voxels
        if weights is not None:
            weights = np.asarray(weights)
            if weights.size != self._data.shape[0]:
                raise ValueError('The size of array with voxel weights/intensities '
                                 'does not match the array of voxel indices!')
            self._data[:, 0] = np.clip(weights, 0.0, 1.0)

    def _set_voxels(self, voxels):
        voxels = np.asarray(voxels)
        if voxels.ndim > 2 or voxels.shape[-1] != 3:
            raise ValueError('Shape of the voxel array must be (n, 3)!')
        if voxels.ndim == 1:
            voxels = voxels.reshape(1, 3)
        if self._data is None or self._data.shape[0] != voxels.shape[0]:
            dtype = self._data.dtype if self._data is not None else None
            self._data = np.ones((voxels.shape[0], 4), dtype=dtype)
        self._data[:, 1:] = voxels

    voxels = property(lambda self: self._data[:, 1:], _set_voxels)

    def _set_weights(self, weights):
        self._data[:, 0] = np.clip(weights, 0.0, 1.0)

    weights = property(lambda self: self._data[:, 0], _set_weights)

    def update(self, other):
        if isinstance(other, IsotropicVoxels):
            self.voxels, self.weights = other.voxels, other.weights
        elif isinstance(other, dict):
            self.voxels = other.get('voxels', self.voxels)
            self.weights = other.get('weights', self.weights)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        for voxel in self.voxels:
            if not mc.voxels.isvalid(voxel):
                raise ValueError('Voxel index ({}, {}, {}) is not valid!'.format(*voxel))
        if self._data.dtype != mc.types.np_float:
            self._data = self._data.astype(mc.types.np_float)
        entry = mc.append_r_lut(np.reshape(self._data, (self._data.size,)))
        target.position.fromarray(mc.voxels.center(np.mean(self.voxels, 0)))
        target.n = self._data.shape[0]
        target.offset = entry.offset
        return target, None, None

    def todict(self):
        return {'voxels': self.voxels.tolist(), 'weights': self.weights.tolist(),
                'type': type(self).__name__}
