"""Sample surface layouts of the layered simulator (mirror of
``xopto/mcml/mcsurface``: base.py, lambertian.py, probe/sixaroundone.py).

A layout customises what a packet sees when it reaches the top (z = 0) or the
bottom surface of the sample: the device handler either overrides the
refractive index / critical cosine of the surrounding medium at the point of
incidence (fibre core, cladding, probe cut-out) and lets the regular Fresnel
logic continue, or reflects the packet itself (Lambertian / specular reflector,
stainless-steel probe tip).  CUDA side: ``csrc/kernels/xo_surface.cuh``.
"""
from typing import Tuple

import numpy as np

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil import boundary, geometry
from ..mcbase.mcutil.fiber import MultimodeFiber, FiberLayout  # noqa: F401

TOP = 'top'
BOTTOM = 'bottom'
NONE = 'none'


class SurfaceLayoutBase(McObject):
    """Base of all layouts (mcsurface/base.py:38-72)."""

    def __init__(self, location: str = NONE):
        super().__init__()
        if location not in (TOP, BOTTOM, NONE):
            raise ValueError('Surface layout location must be "{}", "{}" or "{}"!'.format(
                NONE, TOP, BOTTOM))
        self._location = location

    def _set_location(self, location):
        if location not in (TOP, BOTTOM, NONE):
            raise ValueError('Surface layout location must be "{}", "{}" or "{}"!'.format(
                NONE, TOP, BOTTOM))
        self._location = location

    location = property(lambda self: self._location, _set_location, None,
                        'Location of the surface layout.')


SurfaceLayoutTop = SurfaceLayoutBottom = SurfaceLayoutAny = SurfaceLayoutBase


class SurfaceLayoutDefault(SurfaceLayoutBase):
    """Absent layout: ``{int64 dummy}`` (mcsurface/base.py:124-232)."""
    cu_type = 'xo::SurfNone'

    def cl_type(self, mc):
        class ClSurfaceLayoutDefault(cltypes.Structure):
            _fields_ = [('dummy', cltypes.cl_int64_t)]
        return ClSurfaceLayoutDefault

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.dummy = 0
        return target

    def todict(self):
        return {'type': 'SurfaceLayoutDefault'}

    @classmethod
    def fromdict(cls, data):
        return cls()


class LambertianReflector(SurfaceLayoutBase):
    """Lambertian / specular reflector covering the whole surface
    (mcsurface/lambertian.py): nothing leaves through this surface."""
    cu_type = 'xo::SurfLambertian'

    def cl_type(self, mc):
        T = mc.types
        class ClLambertianReflector(cltypes.Structure):
            _fields_ = [('reflectance', T.mc_fp_t), ('specular', T.mc_fp_t)]
        return ClLambertianReflector

    def __init__(self, reflectance: float = 1.0, specular: float = 0.0):
        super().__init__()
        self.reflectance = reflectance
        self.specular = specular

    def _set_reflectance(self, v):
        self._reflectance = min(max(float(v), 0.0), 1.0)

    reflectance = property(lambda self: self._reflectance, _set_reflectance, None,
                           'Total reflectance of the surface.')

    def _set_specular(self, v):
        self._specular = min(max(float(v), 0.0), 1.0)

    specular = property(lambda self: self._specular, _set_specular, None,
                        'Specular fraction of the reflectance (0: ideal Lambertian).')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.reflectance = self._reflectance
        # Reference quirk kept for parity (lambertian.py:120-135,163): the setter
        # stores the fraction in `_specular_fraction` while cl_pack reads
        # `_specular`, which keeps its initial 0.0 - the specular fraction never
        # reaches the kernel (the reflector is always ideally Lambertian).
        target.specular = 0.0
        return target

    def todict(self):
        return {'type': 'LambertianReflector', 'reflectance': self._reflectance,
                'specular': self._specular}

    def __repr__(self):
        return 'LambertianReflector(reflectance={}, specular={})'.format(
            self._reflectance, self._specular)


class SixAroundOne(SurfaceLayoutBase):
    """Six-around-one fibre probe pressed against the surface
    (mcsurface/probe/sixaroundone.py): fibre cores and claddings, an optional
    filled cut-out that accommodates the fibres, and the reflective probe tip."""
    cu_type = 'xo::SurfSixAroundOne'

    def cl_type(self, mc):
        T = mc.types
        class ClSixAroundOne(cltypes.Structure):
            _fields_ = [
                ('transformation', T.mc_matrix3f_t), ('position', T.mc_point2f_t),
                ('core_spacing', T.mc_fp_t),
                ('cladding_r_squared', T.mc_fp_t), ('cladding_n', T.mc_fp_t),
                ('cladding_cos_critical', T.mc_fp_t),
                ('core_r_squared', T.mc_fp_t), ('core_n', T.mc_fp_t),
                ('core_cos_critical', T.mc_fp_t),
                ('cutout_r_squared', T.mc_fp_t), ('cutout_n', T.mc_fp_t),
                ('cutout_cos_critical', T.mc_fp_t),
                ('probe_r_squared', T.mc_fp_t), ('probe_reflectivity', T.mc_fp_t)]
        return ClSixAroundOne

    def __init__(self, fiber, spacing: float = None, diameter: float = 0.0,
                 reflectivity: float = 1.0, cutout: float = 0.0, cutoutn: float = 1.0,
                 position: Tuple[float, float] = (0.0, 0.0),
                 direction: Tuple[float, float, float] = (0.0, 0.0, 1.0)):
        super().__init__()
        if isinstance(fiber, SixAroundOne):
            o = fiber
            fiber, spacing, cutout, cutoutn = o.fiber, o.spacing, o._cutout, o.cutoutn
            reflectivity, diameter = o.reflectivity, o.diameter
            position, direction = o.position, o.direction
        elif spacing is None:
            spacing = fiber.dcladding
        self._fiber = fiber
        self._position = np.zeros((2,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self.spacing = spacing
        self.diameter = diameter
        self.reflectivity = reflectivity
        self.cutout = cutout
        self.cutoutn = cutoutn
        self.position = position
        self.direction = direction

    def _set_fiber(self, fiber):
        self._fiber = fiber

    fiber = property(lambda self: self._fiber, _set_fiber, None, 'Multimode optical fiber.')

    def _set_spacing(self, v):
        self._spacing = max(float(v), 0.0)

    spacing = property(lambda self: self._spacing, _set_spacing, None,
                       'Spacing of the optical fibers (m).')

    def _set_diameter(self, v):
        self._diameter = max(float(v), 0.0)

    diameter = property(lambda self: self._diameter, _set_diameter, None,
                        'Outer diameter of the probe tip (m).')

    def _set_reflectivity(self, v):
        self._reflectivity = min(max(float(v), 0.0), 1.0)

    reflectivity = property(lambda self: self._reflectivity, _set_reflectivity, None,
                            'Reflectivity of the probe tip.')

    def _set_cutout(self, v):
        self._cutout = max(float(v), 0.0)

    cutout = property(lambda self: self._cutout, _set_cutout, None,
                      'Diameter of the cut-out accommodating the fibers (0: none).')

    def _set_cutoutn(self, v):
        self._cutout_n = max(float(v), 1.0)

    cutoutn = property(lambda self: self._cutout_n, _set_cutoutn, None,
                       'Refractive index of the cut-out fill.')

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position, None,
                        'Position of the probe centre (x, y).')

    def _set_direction(self, d):
        self._direction[:] = d
        norm = np.linalg.norm(self._direction)
        if norm == 0.0:
            raise ValueError('Direction vector norm/length must not be 0!')
        self._direction *= 1.0/norm

    direction = property(lambda self: self._direction, _set_direction, None,
                         'Direction of the optical fibers.')

    def check(self) -> bool:
        if self._fiber.dcladding > self._spacing:
            raise ValueError('The optical fibers are overlapping!')
        if self._diameter < self._fiber.dcladding:
            raise ValueError('The probe diameter is too small to accommodate the fibers')
        if self._cutout != 0.0 and self._cutout < self._spacing + self._fiber.dcladding:
            raise ValueError('The cutout is too small to accommodate the optical fibers!')
        return True

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        adir = self._direction[0], self._direction[1], abs(self._direction[2])
        n_sample = mc.layers[1].n if self.location == TOP else mc.layers[-2].n
        target.transformation.fromarray(geometry.transform_base(adir, (0.0, 0.0, 1.0)))
        target.position.fromarray(self._position)
        target.core_spacing = self._spacing
        target.core_r_squared = 0.25*self._fiber.dcore**2
        target.core_n = self._fiber.ncore
        target.core_cos_critical = boundary.cos_critical(n_sample, self._fiber.ncore)
        target.cladding_r_squared = 0.25*self._fiber.dcladding**2
        target.cladding_n = self._fiber.ncladding
        target.cladding_cos_critical = boundary.cos_critical(n_sample, self._fiber.ncladding)
        target.cutout_r_squared = 0.25*self._cutout**2
        target.cutout_n = self._cutout_n
        target.cutout_cos_critical = boundary.cos_critical(n_sample, self._cutout_n)
        target.probe_r_squared = 0.25*self._diameter**2
        target.probe_reflectivity = self._reflectivity
        return target

    def todict(self):
        return {'type': 'SixAroundOne', 'fiber': self._fiber.todict(),
                'spacing': self._spacing, 'cutout': self._cutout, 'cutoutn': self._cutout_n,
                'reflectivity': self._reflectivity, 'diameter': self._diameter,
                'position': self._position.tolist(), 'direction': self._direction.tolist()}

    def __repr__(self):
        return 'SixAroundOne(fiber={}, spacing={}, diameter={}, reflectivity={}, cutout={}, ' \
               'cutoutn={})'.format(self._fiber, self._spacing, self._diameter,
                                    self._reflectivity, self._cutout, self._cutout_n)


class LinearArray(SurfaceLayoutBase):
    """Linear array of ``n`` fibers in a stainless-steel probe pressed against the
    surface (mcsurface/probe/lineararray.py): cores, claddings, an optional
    rectangular filled cut-out and the reflective probe tip.  ``n`` is a
    compile-time feature."""
    def cu_type(self, mc):
        return 'xo::SurfLinearArray<{}>'.format(self._n)

    def cl_type(self, mc):
        T = mc.types
        class ClLinearArray(cltypes.Structure):
            _fields_ = [
                ('transformation', T.mc_matrix3f_t),
                ('cutout_transformation', T.mc_matrix2f_t),
                ('position', T.mc_point2f_t), ('first_position', T.mc_point2f_t),
                ('delta_position', T.mc_point2f_t), ('core_spacing', T.mc_fp_t),
                ('cladding_r_squared', T.mc_fp_t), ('cladding_n', T.mc_fp_t),
                ('cladding_cos_critical', T.mc_fp_t),
                ('core_r_squared', T.mc_fp_t), ('core_n', T.mc_fp_t),
                ('core_cos_critical', T.mc_fp_t),
                ('cutout_width_half', T.mc_fp_t), ('cutout_height_half', T.mc_fp_t),
                ('cutout_n', T.mc_fp_t), ('cutout_cos_critical', T.mc_fp_t),
                ('probe_r_squared', T.mc_fp_t), ('probe_reflectivity', T.mc_fp_t)]
        return ClLinearArray

    def __init__(self, fiber, n: int = 1, spacing: float = None,
                 orientation: Tuple[float, float] = (1.0, 0.0), diameter: float = 0.0,
                 reflectivity: float = 1.0, cutout: Tuple[float, float] = (0.0, 0.0),
                 cutoutn: float = 1.0, position: Tuple[float, float] = (0.0, 0.0),
                 direction: Tuple[float, float, float] = (0.0, 0.0, 1.0)):
        super().__init__()
        if isinstance(fiber, LinearArray):
            o = fiber
            fiber, n, spacing, orientation = o.fiber, o.n, o.spacing, o.orientation
            diameter, reflectivity, cutout, cutoutn = o.diameter, o.reflectivity, o.cutout, o.cutoutn
            position, direction = o.position, o.direction
        elif spacing is None:
            spacing = fiber.dcladding
        self._fiber = fiber
        self._n = max(int(n), 1)
        self._cutout = np.zeros((2,))
        self._orientation = np.array((1.0, 0.0))
        self._position = np.zeros((2,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self.spacing, self.diameter, self.reflectivity = spacing, diameter, reflectivity
        self.cutout, self.cutoutn = cutout, cutoutn
        self.orientation, self.position, self.direction = orientation, position, direction

    def _set_fiber(self, fiber):
        self._fiber = fiber

    fiber = property(lambda self: self._fiber, _set_fiber)
    n = property(lambda self: self._n)

    def _set_spacing(self, v):
        self._spacing = float(v)

    spacing = property(lambda self: self._spacing, _set_spacing)

    def _set_diameter(self, v):
        self._diameter = max(float(v), 0.0)

    diameter = property(lambda self: self._diameter, _set_diameter)

    def _set_reflectivity(self, v):
        self._reflectivity = min(max(float(v), 0.0), 1.0)

    reflectivity = property(lambda self: self._reflectivity, _set_reflectivity)

    def _set_orientation(self, o):
        self._orientation[:] = o
        norm = np.linalg.norm(self._orientation)
        if norm == 0.0:
            raise ValueError('Orientation vector norm/length must not be 0!')
        self._orientation *= 1.0/norm

    orientation = property(lambda self: self._orientation, _set_orientation)

    def _set_cutout(self, c):
        self._cutout[:] = np.maximum(0.0, c)

    cutout = property(lambda self: self._cutout, _set_cutout, None,
                      'Size (width, height) of the cut-out that accommodates the fibers.')

    def _set_cutoutn(self, v):
        self._cutout_n = max(float(v), 1.0)

    cutoutn = property(lambda self: self._cutout_n, _set_cutoutn)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def _set_direction(self, d):
        self._direction[:] = d
        norm = np.linalg.norm(self._direction)
        if norm == 0.0:
            raise ValueError('Direction vector norm/length must not be 0!')
        self._direction *= 1.0/norm

    direction = property(lambda self: self._direction, _set_direction)

    def fiber_position(self, index: int) -> Tuple[float, float]:
        if index >= self._n or index < -self._n:
            raise IndexError('The fiber index is out of valid range!')
        left = self._position - self._orientation*self._spacing*(self._n - 1)*0.5
        return tuple(left + self._spacing*self._orientation*int(index))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        adir = self._direction[0], self._direction[1], abs(self._direction[2])
        n_sample = mc.layers[1].n if self.location == TOP else mc.layers[-2].n
        target.transformation.fromarray(geometry.transform_base(adir, (0.0, 0.0, 1.0)))
        target.position.fromarray(self._position)
        target.first_position.fromarray(self.fiber_position(0))
        target.delta_position.fromarray(self._orientation*self._spacing)
        target.core_r_squared = 0.25*self._fiber.dcore**2
        target.core_n = self._fiber.ncore
        target.core_cos_critical = boundary.cos_critical(n_sample, self._fiber.ncore)
        target.cladding_r_squared = 0.25*self._fiber.dcladding**2
        target.cladding_n = self._fiber.ncladding
        target.cladding_cos_critical = boundary.cos_critical(n_sample, self._fiber.ncladding)
        target.probe_r_squared = 0.25*self._diameter**2
        target.probe_reflectivity = self._reflectivity
        target.cutout_transformation.fromarray(
            geometry.rotation_matrix_2d(self._orientation, [1.0, 0.0]))
        target.cutout_width_half = self._cutout[0]*0.5
        target.cutout_height_half = self._cutout[1]*0.5
        target.cutout_n = self._cutout_n
        target.cutout_cos_critical = boundary.cos_critical(n_sample, self._cutout_n)
        return target

    def todict(self):
        return {'type': 'LinearArray', 'fiber': self._fiber.todict(), 'n': self._n,
                'spacing': self._spacing, 'orientation': self._orientation.tolist(),
                'diameter': self._diameter, 'reflectivity': self._reflectivity,
                'cutout': self._cutout.tolist(), 'cutoutn': self._cutout_n,
                'position': self._position.tolist(), 'direction': self._direction.tolist()}

    def __repr__(self):
        return 'LinearArray(fiber={}, n={}, spacing={}, diameter={})'.format(
            self._fiber, self._n, self._spacing, self._diameter)


class FiberArray(SurfaceLayoutBase):
    """Individually placed / tilted fibers in a stainless-steel probe
    (mcsurface/probe/fiberarray.py).  The number of fibers is a compile-time
    feature.  Quirk kept: the reference packs the tip reflectivity into an
    attribute that is not a struct field (``target.probe_reflectivity`` while the
    field is called ``reflectivity``, fiberarray.py:81,381), so the kernel always
    sees 0 - a packet that hits the tip loses all its weight."""
    def cu_type(self, mc):
        return 'xo::SurfFiberArray<{}>'.format(len(self._fibers))

    def cl_type(self, mc):
        T = mc.types
        n = self.n
        class ClFiberArray(cltypes.Structure):
            _fields_ = [
                ('transformation', T.mc_matrix3f_t*n), ('fiber_position', T.mc_point2f_t*n),
                ('cladding_r_squared', T.mc_fp_t*n), ('cladding_n', T.mc_fp_t*n),
                ('cladding_cos_critical', T.mc_fp_t*n),
                ('core_r_squared', T.mc_fp_t*n), ('core_n', T.mc_fp_t*n),
                ('core_cos_critical', T.mc_fp_t*n),
                ('probe_position', T.mc_point2f_t), ('probe_r_squared', T.mc_fp_t),
                ('reflectivity', T.mc_fp_t)]
        return ClFiberArray

    def __init__(self, fibers, diameter: float = 0.0, reflectivity: float = 0.0,
                 position: Tuple[float, float] = (0.0, 0.0)):
        super().__init__()
        if isinstance(fibers, FiberArray):
            o = fibers
            fibers, diameter, reflectivity, position = o.fibers, o.diameter, o.reflectivity, o.position
        self._fibers = list(fibers)
        self._position = np.zeros((2,))
        self.diameter, self.reflectivity, self.position = diameter, reflectivity, position

    def _set_fibers(self, fibers):
        if len(self._fibers) != len(fibers):
            raise ValueError('The number of optical fibers must not change!')
        self._fibers[:] = fibers

    fibers = property(lambda self: self._fibers, _set_fibers)
    n = property(lambda self: len(self._fibers))

    def _set_diameter(self, v):
        self._diameter = max(float(v), 0.0)

    diameter = property(lambda self: self._diameter, _set_diameter)

    def _set_reflectivity(self, v):
        self._reflectivity = min(max(float(v), 0.0), 1.0)

    reflectivity = property(lambda self: self._reflectivity, _set_reflectivity)

    def _set_position(self, p):
        self._position[:] = p

    position = property(lambda self: self._position, _set_position)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        n_sample = mc.layers[1].n if self.location == TOP else mc.layers[-2].n
        for index, cfg in enumerate(self._fibers):
            adir = cfg.direction[0], cfg.direction[1], abs(cfg.direction[2])
            target.transformation[index].fromarray(
                geometry.transform_base(adir, (0.0, 0.0, 1.0)))
            target.fiber_position[index].fromarray(cfg.position)
            target.cladding_r_squared[index] = 0.25*cfg.fiber.dcladding**2
            target.cladding_n[index] = cfg.fiber.ncladding
            target.cladding_cos_critical[index] = boundary.cos_critical(
                n_sample, cfg.fiber.ncladding)
            target.core_r_squared[index] = 0.25*cfg.fiber.dcore**2
            target.core_n[index] = cfg.fiber.ncore
            target.core_cos_critical[index] = boundary.cos_critical(n_sample, cfg.fiber.ncore)
        target.probe_position.fromarray(self._position)
        target.probe_r_squared = 0.25*self._diameter**2
        # (the struct field `reflectivity` stays 0, see the class docstring)
        return target

    def todict(self):
        return {'type': 'FiberArray', 'fibers': [f.todict() for f in self._fibers],
                'diameter': self._diameter, 'reflectivity': self._reflectivity,
                'position': self._position.tolist()}

    def __repr__(self):
        return 'FiberArray(fibers={}, diameter={})'.format(self._fibers, self._diameter)


class SurfaceLayouts(McObject):
    """Container {top, bottom} (mcsurface/base.py:235-426)."""

    def __init__(self, top=None, bottom=None):
        super().__init__()
        if isinstance(top, SurfaceLayouts):
            sl = top
            top, bottom = sl.top, sl.bottom
        top = SurfaceLayoutDefault() if top is None else top
        bottom = SurfaceLayoutDefault() if bottom is None else bottom
        if top.location not in (NONE, TOP) and not isinstance(top, SurfaceLayoutDefault):
            raise ValueError('The top surface layout is already assigned to another surface!')
        if bottom.location not in (NONE, BOTTOM) and \
                not isinstance(bottom, SurfaceLayoutDefault):
            raise ValueError('The bottom surface layout is already assigned to another '
                             'surface!')
        top.location, bottom.location = TOP, BOTTOM
        self._top, self._bottom = top, bottom

    top = property(lambda self: self._top)
    bottom = property(lambda self: self._bottom)

    def cl_type(self, mc):
        class ClSurfaceLayouts(cltypes.Structure):
            _fields_ = [('top', self._top.fetch_cl_type(mc)),
                        ('bottom', self._bottom.fetch_cl_type(mc))]
        return ClSurfaceLayouts

    def cl_options(self, mc):
        out, used = [], False
        if type(self._top) is not SurfaceLayoutDefault:
            out.append(('MC_USE_TOP_SURFACE_LAYOUT', True))
            out.extend(self._top.fetch_cl_options(mc))
            used = True
        if type(self._bottom) is not SurfaceLayoutDefault:
            out.append(('MC_USE_BOTTOM_SURFACE_LAYOUT', True))
            out.extend(self._bottom.fetch_cl_options(mc))
            used = True
        if used:
            out.append(('MC_USE_SURFACE_LAYOUTS', True))
        return out

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        self._top.cl_pack(mc, target.top)
        self._bottom.cl_pack(mc, target.bottom)
        return target

    def types(self):
        return type(self._top), type(self._bottom)

    def todict(self):
        return {'type': 'SurfaceLayouts', 'top': self._top.todict(),
                'bottom': self._bottom.todict()}

    def __iter__(self):
        return iter([self._top, self._bottom])

    def __repr__(self):
        return 'SurfaceLayouts(top={}, bottom={})'.format(self._top, self._bottom)
