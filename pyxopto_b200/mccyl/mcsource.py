"""Photon packet sources of the cylindrical simulator (mirror of
``xopto/mccyl/mcsource``: Line, GaussianBeam, UniformBeam, IsotropicPoint).
Beams propagate along +x by default and enter the sample through the outer
cylinder surface."""
import numpy as np

from ..cl import cltypes
from ..mcbase.mcutil import boundary, geometry
from ..mcml.mcsource import Source, _unit


class _Positioned(Source):
    def _set_position(self, p):
        self._position[:] = p

    def _set_direction(self, d):
        self._direction[:] = _unit(d)

    position = property(lambda self: self._position, _set_position, None,
                        'Source position.')
    direction = property(lambda self: self._direction, _set_direction, None,
                         'Source direction.')


class Line(_Positioned):
    """Infinitely thin beam (mccyl/mcsource/line.py)."""
    cu_type = 'xo::CylSrcLine'
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

    def __init__(self, position=(0.0, 0.0, 0.0), direction=(1.0, 0.0, 0.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.zeros((3,))
        self.position = position
        self.direction = direction

    def cl_pack(self, mc, target=None):
        """The line is propagated (forwards or backwards) to its entry point on
        the sample surface and refracted there (line.py:181-236; the surface
        normal handed to the Fresnel term is not normalised in the reference,
        kept)."""
        if target is None:
            target = self.cl_type(mc)()
        intersection, normal = mc.layers.intersect(
            self._position, self._direction, entrance=True)
        if intersection is None:
            raise ValueError('The Line source does not intersect the sample!')
        costheta = np.dot(normal, self._direction)
        reflectance = boundary.reflectance(mc.layer(0).n, mc.layer(1).n, costheta)
        if reflectance >= 1.0:
            raise ValueError('The line source is fully reflected from the '
                             'sample surface!')
        refracted = boundary.refract(self._direction, normal, mc.layer(0).n, mc.layer(1).n)
        reflected = boundary.reflect(self._direction, normal)
        target.position.fromarray(intersection)
        target.direction_medium.fromarray(self._direction)
        target.direction_sample.fromarray(refracted)
        target.direction_reflected.fromarray(reflected)
        target.reflectance = reflectance
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': 'Line'}


class GaussianBeam(_Positioned):
    """Collimated Gaussian beam (mccyl/mcsource/gaussianbeam.py)."""
    cu_type = 'xo::CylSrcGaussianBeam'
    _update_keys = ('sigma', 'clip', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClGaussianBeam(cltypes.Structure):
            _pack_ = 1
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('sigma', T.mc_point2f_t),
                        ('clip', T.mc_fp_t)]
        return ClGaussianBeam

    @staticmethod
    def fwhm2sigma(fwhm: float) -> float:
        return fwhm/(8*np.log(2))**0.5

    @staticmethod
    def sigma2fwhm(sigma: float) -> float:
        return sigma*(8*np.log(2))**0.5

    def __init__(self, sigma, clip: float = 5.0, position=(0.0, 0.0, 0.0),
                 direction=(1.0, 0.0, 0.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.zeros((3,))
        self._sigma = np.zeros((2,))
        self.sigma, self.clip = sigma, clip
        self.position, self.direction = position, direction

    def _set_sigma(self, s):
        self._sigma[:] = s
        if np.any(self._sigma < 0.0):
            raise ValueError('Beam diameter/sigma must not be negative!')

    def _set_clip(self, c):
        self._clip = float(c)
        if self._clip < 0.0:
            raise ValueError('Clip diameter/sigma must be greater than zero!.')

    sigma = property(lambda self: self._sigma, _set_sigma)
    clip = property(lambda self: self._clip, _set_clip)

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
                'direction': self._direction.tolist(), 'type': 'GaussianBeam'}


class UniformBeam(_Positioned):
    """Collimated elliptical top-hat beam (mccyl/mcsource/uniformbeam.py)."""
    cu_type = 'xo::CylSrcUniformBeam'
    _update_keys = ('diameter', 'position', 'direction')

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClUniformBeam(cltypes.Structure):
            _fields_ = [('transformation', T.mc_matrix3f_t), ('position', T.mc_point3f_t),
                        ('direction', T.mc_point3f_t), ('radius', T.mc_point2f_t)]
        return ClUniformBeam

    def __init__(self, diameter, position=(0.0, 0.0, 0.0), direction=(1.0, 0.0, 0.0)):
        super().__init__()
        self._position = np.zeros((3,))
        self._direction = np.zeros((3,))
        self._diameter = np.zeros((2,))
        self.diameter = diameter
        self.position, self.direction = position, direction

    def _set_diameter(self, d):
        self._diameter[:] = d
        self._diameter = np.maximum(0.0, self._diameter)

    diameter = property(lambda self: self._diameter, _set_diameter)

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
        return {'diameter': self._diameter.tolist(),
                'position': self._position.tolist(),
                'direction': self._direction.tolist(), 'type': 'UniformBeam'}


class IsotropicPoint(Source):
    """Isotropic point source inside or outside of the sample
    (mccyl/mcsource/point.py)."""
    cu_type = 'xo::CylSrcIsotropicPoint'
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
        target.position.fromarray(self._position)
        r = (self._position[0]**2 + self._position[1]**2)**0.5
        if r >= mc.layers.diameter()*0.5:
            target.layer_index = 1
        else:
            target.layer_index = mc.layer_index(r)
        return target, None, None

    def todict(self):
        return {'position': self._position.tolist(), 'type': 'IsotropicPoint'}
