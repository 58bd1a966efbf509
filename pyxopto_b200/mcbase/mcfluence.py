"""Fluence / deposition accumulators (mirror of ``xopto/mcbase/mcfluence``:
Fluence, FluenceRz, Fluencet).  ``mode='deposition'`` accumulates absorbed
weight, ``mode='fluence'`` divides each deposit by the local mua."""
from typing import Tuple

import numpy as np

from ..cl import cltypes
from .mcobject import McObject
from .mcutil.axis import Axis, RadialAxis  # noqa: F401  (re-exported like the reference)

_K_DEFAULT = 0x7FFFFF


def _check_mode(mode):
    if mode not in ('fluence', 'deposition'):
        raise ValueError('The value of mode parameter must be '
                         '"fluence" or "deposition" but got {}!'.format(mode))


class _FluenceBase(McObject):
    def _init_common(self, mode, k=_K_DEFAULT, data=None, nphotons=0):
        _check_mode(mode)
        self._mode, self._k, self._data, self._nphotons = mode, k, data, nphotons

    nphotons = property(lambda self: self._nphotons)
    mode = property(lambda self: self._mode)

    # Device-resident accumulation (``Mc.lazy_fluence``): the simulator keeps the float64
    # grid of this result on the device across ``run(out=...)`` calls and hands over a
    # loader; the grid comes to the host when ``raw`` / ``data`` (anything that touches
    # ``_data``) is read.
    _pending = None
    _store = None

    def _materialize(self):
        loader, self._pending = self._pending, None
        if loader is not None:
            loader(self)

    def _get_store(self):
        self._materialize()
        return self._store

    def _set_store(self, data):
        self._materialize()
        self._store = data

    _data = property(_get_store, _set_store)

    def _set_k(self, k):
        self._k = max(1, min(int(k), int(2**31 - 1)))

    k = property(lambda self: self._k, _set_k, None,
                 'Weight to fixed-point conversion factor.')

    def _set_raw(self, data):
        self._data = data

    raw = property(lambda self: self._data, _set_raw, None, 'Raw accumulator data.')

    def cl_options(self, mc):
        return [('MC_USE_FLUENCE', True),
                ('MC_FLUENCE_MODE_RATE', self._mode == 'fluence')] + self._extra_options

    _extra_options = []

    def update_data(self, mc, data, nphotons, **kwargs):
        accumulators = data[np.dtype(mc.types.np_accu)]
        if self._data is not None:
            self._data += np.reshape(accumulators[0], self._data.shape)*(1.0/self.k)
            self._nphotons += nphotons
        else:
            self._data = accumulators[0]*(1.0/self.k)
            self._data.shape = self.shape
            self._nphotons = nphotons

    def update_scaled(self, scaled: np.ndarray, nphotons: int, owned: bool = False):
        """``update_data`` with the conversion ``accumulators*(1/k)`` already done
        (bit-identically) on the device: ``scaled`` is a float64 array that the
        caller may reuse afterwards - unless ``owned``: then a fresh result keeps the
        array itself (no copy of a 65 MB grid)."""
        if self._data is not None:
            self._data += np.reshape(scaled, self._data.shape)
            self._nphotons += nphotons
        else:
            self._data = np.reshape(scaled, self.shape) if owned else \
                np.array(scaled, dtype=np.float64).reshape(self.shape)
            self._nphotons = nphotons

    def cu_window(self, mc, max_bins: int, focus):
        """(org0, org1, org2, ext0, ext1, ext2): block of grid cells around the
        point ``focus`` (source position) that each CTA accumulates in shared
        memory (xo::FluWindow); at most ``max_bins`` cells.  Default: none."""
        return (0, 0, 0, 0, 0, 0)

    @staticmethod
    def _window_axis(axis, n_want: int, at: float, lead: float = 0.25):
        """First cell and cell count of a window of ~n_want cells of ``axis``
        placed so that ``at`` lies ``lead`` of the way into it."""
        n = int(min(max(n_want, 1), axis.n))
        i_at = int(np.floor((at - axis.start)/axis.step)) if axis.step != 0.0 else 0
        first = int(min(max(i_at - int(lead*n), 0), axis.n - n))
        return first, n

    def update(self, obj):
        if self._data is not None:
            if self.shape != obj.shape:
                raise TypeError('Cannot update with fluence data of incompatible shape!')
            self._data += obj.raw
            self._nphotons += obj.nphotons
        else:
            self._data = obj.raw
            self._nphotons = obj.nphotons


class Fluence(_FluenceBase):
    """Cartesian x-y-z grid; raw data shape (nz, ny, nx) (mcfluence/fluence.py)."""
    cu_type = 'xo::FluXyz'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluence(cltypes.Structure):
            _fields_ = [('inv_step', T.mc_point3f_t), ('top_left', T.mc_point3f_t),
                        ('shape', T.mc_point3s_t), ('offset', T.mc_size_t),
                        ('k', T.mc_int_t)]
        return ClFluence

    def __init__(self, xaxis=None, yaxis: Axis = None, zaxis: Axis = None,
                 mode: str = 'deposition'):
        super().__init__()
        data, nphotons, k = None, 0, _K_DEFAULT
        if isinstance(xaxis, Fluence):
            f = xaxis
            xaxis, yaxis, zaxis = Axis(f.xaxis), Axis(f.yaxis), Axis(f.zaxis)
            nphotons, mode, k = f.nphotons, f.mode, f.k
            if f.raw is not None:
                data = np.copy(f.raw)
        xaxis = Axis(-0.5, 0.5, 1) if xaxis is None else xaxis
        yaxis = Axis(-0.5, 0.5, 1) if yaxis is None else yaxis
        zaxis = Axis(0.0, 1.0, 1) if zaxis is None else zaxis
        for name, ax in (('x', xaxis), ('y', yaxis), ('z', zaxis)):
            if ax.logscale:
                raise ValueError('Fluence does not support logarithmic {} axis!'.format(name))
        self._x_axis, self._y_axis, self._z_axis = xaxis, yaxis, zaxis
        self._init_common(mode, k, data, nphotons)

    shape = property(lambda self: (self._z_axis.n, self._y_axis.n, self._x_axis.n))
    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    zaxis = property(lambda self: self._z_axis)
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    dx = property(lambda self: abs(self._x_axis.step))
    dy = property(lambda self: abs(self._y_axis.step))
    dz = property(lambda self: abs(self._z_axis.step))

    @property
    def data(self):
        return self._data*(1.0/(max(self.nphotons, 1)*self.dx*self.dy*self.dz))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        for c, ax in zip('xyz', (self._x_axis, self._y_axis, self._z_axis)):
            setattr(target.top_left, c, ax.start)
            setattr(target.inv_step, c, 1.0/ax.step)
            setattr(target.shape, c, ax.n)
        target.k = self._k
        return target

    def cu_window(self, mc, max_bins: int, focus):
        axes = (self._x_axis, self._y_axis, self._z_axis)
        steps = [abs(a.step) for a in axes]
        if max_bins < 8 or min(steps) <= 0.0:
            return (0, 0, 0, 0, 0, 0)
        # cube of equal physical edge, clipped to the grid; spare cells go to z
        edge = (max_bins*steps[0]*steps[1]*steps[2])**(1.0/3.0)
        nx = int(min(max(edge//steps[0], 1), axes[0].n))
        ny = int(min(max(edge//steps[1], 1), axes[1].n))
        nz = int(min(max(max_bins//(nx*ny), 1), axes[2].n))
        x0, nx = self._window_axis(axes[0], nx, focus[0], 0.5)
        y0, ny = self._window_axis(axes[1], ny, focus[1], 0.5)
        z0, nz = self._window_axis(axes[2], nz, focus[2], 0.25)
        return (x0, y0, z0, nx, ny, nz)

    def todict(self):
        return {'type': 'Fluence', 'mode': self._mode, 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'zaxis': self._z_axis.todict()}

    @classmethod
    def fromdict(cls, data):
        d = dict(data)
        d.pop('type')
        return cls(Axis.fromdict(d.pop('xaxis')), Axis.fromdict(d.pop('yaxis')),
                   Axis.fromdict(d.pop('zaxis')), **d)


class FluenceRz(_FluenceBase):
    """Radially symmetric r-z grid (mcfluence/fluencerz.py).

    Quirk kept from the reference: ``shape`` is reported as (n_r, n_z) although
    the kernel stores bins z-major (index = iz*n_r + ir, fluencerz.py:140,292).
    ``raw_zr`` exposes the data with its true (n_z, n_r) layout."""
    cu_type = 'xo::FluRz'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluenceRz(cltypes.Structure):
            _fields_ = [('center', T.mc_point3f_t), ('inv_dr', T.mc_fp_t),
                        ('inv_dz', T.mc_fp_t), ('n_r', T.mc_size_t),
                        ('n_z', T.mc_size_t), ('offset', T.mc_size_t),
                        ('k', T.mc_int_t)]
        return ClFluenceRz

    def __init__(self, raxis=None, zaxis: Axis = None,
                 center: Tuple[float, float] = (0.0, 0.0), mode: str = 'deposition'):
        super().__init__()
        data, nphotons, k = None, 0, _K_DEFAULT
        if isinstance(raxis, FluenceRz):
            f = raxis
            raxis, zaxis, center = Axis(f.raxis), Axis(f.zaxis), f.center
            nphotons, mode, k = f.nphotons, f.mode, f.k
            if f.raw is not None:
                data = np.copy(f.raw)
        raxis = Axis(0.0, 1.0, 1) if raxis is None else raxis
        zaxis = Axis(0.0, 1.0, 1) if zaxis is None else zaxis
        if raxis.logscale or zaxis.logscale:
            raise ValueError('FluenceRz does not support logarithmic axes!')
        self._r_axis, self._z_axis = raxis, zaxis
        self._center = np.zeros((2,))
        self._center[:] = center
        self._init_common(mode, k, data, nphotons)

    shape = property(lambda self: (self._r_axis.n, self._z_axis.n))
    raxis = property(lambda self: self._r_axis)
    zaxis = property(lambda self: self._z_axis)
    r = property(lambda self: self._r_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    dr = property(lambda self: abs(self._r_axis.step))
    dz = property(lambda self: abs(self._z_axis.step))

    def _set_center(self, c):
        self._center[:] = c

    center = property(lambda self: self._center, _set_center)

    @property
    def raw_zr(self):
        return None if self._data is None else \
            self._data.reshape(self._z_axis.n, self._r_axis.n)

    @property
    def data(self):
        area = np.pi*(self._r_axis.edges[1:]**2 - self._r_axis.edges[:-1]**2)
        k = 1.0/(self.nphotons*area*self.dz)
        k.shape = (1, k.size)
        return self._data*k

    @property
    def data_zr(self):
        area = np.pi*(self._r_axis.edges[1:]**2 - self._r_axis.edges[:-1]**2)
        return self.raw_zr/(max(self.nphotons, 1)*area[None, :]*self.dz)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.center.x, target.center.y = self._center
        target.center.z = self._z_axis.start
        target.inv_dr = 1.0/self._r_axis.step if self._r_axis.step != 0.0 else 0.0
        target.inv_dz = 1.0/self._z_axis.step
        target.n_r, target.n_z = self._r_axis.n, self._z_axis.n
        target.k = self._k
        return target

    def cu_window(self, mc, max_bins: int, focus):
        dr, dz = abs(self._r_axis.step), abs(self._z_axis.step)
        if max_bins < 4 or dr <= 0.0 or dz <= 0.0:
            return (0, 0, 0, 0, 0, 0)
        # physical window R x aspect*R (radius x depth), clipped to the grid; light
        # diffuses as far sideways as down: aspect 1 (C2: 0.85-1.0 measured best, 2.0 is 3 % slower)
        aspect = float(getattr(mc, 'fluence_window_aspect', None) or 1.0)
        radius = np.sqrt(max_bins*dr*dz/aspect)
        n_r = int(min(max(radius//dr, 1), self._r_axis.n))
        n_z = int(min(max(max_bins//n_r, 1), self._z_axis.n))
        n_r = int(min(max(max_bins//n_z, 1), self._r_axis.n))
        rc = float(np.hypot(focus[0] - self._center[0], focus[1] - self._center[1]))
        r0, n_r = self._window_axis(self._r_axis, n_r, rc, 0.5)
        z0, n_z = self._window_axis(self._z_axis, n_z, focus[2], 0.25)
        return (r0, z0, 0, n_r, n_z, 1)

    def todict(self):
        return {'type': 'FluenceRz', 'mode': self._mode, 'raxis': self._r_axis.todict(),
                'zaxis': self._z_axis.todict(), 'center': self._center.tolist()}

    @classmethod
    def fromdict(cls, data):
        d = dict(data)
        d.pop('type')
        return cls(Axis.fromdict(d.pop('raxis')), Axis.fromdict(d.pop('zaxis')), **d)


class FluenceRzt(_FluenceBase):
    """Time-resolved radially symmetric grid; bins indexed (z, r, t) in memory,
    ``shape`` reported as (n_r, n_z, n_t) like the reference (mcfluence/fluencerzt.py)."""
    cu_type = 'xo::FluRzt'
    _extra_options = [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluenceRzt(cltypes.Structure):
            _fields_ = [('center', T.mc_point3f_t), ('t_min', T.mc_fp_t),
                        ('inv_dr', T.mc_fp_t), ('inv_dz', T.mc_fp_t), ('inv_dt', T.mc_fp_t),
                        ('n_r', T.mc_size_t), ('n_z', T.mc_size_t), ('n_t', T.mc_size_t),
                        ('offset', T.mc_size_t), ('k', T.mc_int_t)]
        return ClFluenceRzt

    def __init__(self, raxis=None, zaxis: Axis = None, taxis: Axis = None,
                 center: Tuple[float, float] = (0.0, 0.0), mode: str = 'deposition'):
        super().__init__()
        data, nphotons, k = None, 0, _K_DEFAULT
        if isinstance(raxis, FluenceRzt):
            f = raxis
            raxis, zaxis, taxis, center = Axis(f.raxis), Axis(f.zaxis), Axis(f.taxis), f.center
            nphotons, mode, k = f.nphotons, f.mode, f.k
            if f.raw is not None:
                data = np.copy(f.raw)
        raxis = Axis(0.0, 1.0, 1) if raxis is None else raxis
        zaxis = Axis(0.0, 1.0, 1) if zaxis is None else zaxis
        taxis = Axis(0.0, 1.0, 1) if taxis is None else taxis
        if raxis.logscale or zaxis.logscale or taxis.logscale:
            raise ValueError('FluenceRzt does not support logarithmic axes!')
        self._r_axis, self._z_axis, self._t_axis = raxis, zaxis, taxis
        self._center = np.zeros((2,))
        self._center[:] = center
        self._init_common(mode, k, data, nphotons)

    shape = property(lambda self: (self._r_axis.n, self._z_axis.n, self._t_axis.n))
    raxis = property(lambda self: self._r_axis)
    zaxis = property(lambda self: self._z_axis)
    taxis = property(lambda self: self._t_axis)
    r = property(lambda self: self._r_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    t = property(lambda self: self._t_axis.centers)
    center = property(lambda self: self._center)

    @property
    def raw_zrt(self):
        return None if self._data is None else \
            self._data.reshape(self._z_axis.n, self._r_axis.n, self._t_axis.n)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.center.x, target.center.y = self._center
        target.center.z = self._z_axis.start
        target.t_min = self._t_axis.start
        target.inv_dr = 1.0/self._r_axis.step if self._r_axis.step != 0.0 else 0.0
        target.inv_dz = 1.0/self._z_axis.step
        target.inv_dt = 1.0/self._t_axis.step
        target.n_r, target.n_z, target.n_t = self._r_axis.n, self._z_axis.n, self._t_axis.n
        target.k = self._k
        return target

    def todict(self):
        return {'type': 'FluenceRzt', 'mode': self._mode, 'raxis': self._r_axis.todict(),
                'zaxis': self._z_axis.todict(), 'taxis': self._t_axis.todict(),
                'center': self._center.tolist()}


class FluenceCyl(_FluenceBase):
    """Cylindrical r-fi-z grid; raw shape (n_z, n_fi, n_r) (mcfluence/fluencecyl.py)."""
    cu_type = 'xo::FluCyl'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluenceCyl(cltypes.Structure):
            _fields_ = [('center', T.mc_point2f_t), ('r_min', T.mc_fp_t),
                        ('fi_min', T.mc_fp_t), ('z_min', T.mc_fp_t),
                        ('inv_dr', T.mc_fp_t), ('inv_dfi', T.mc_fp_t), ('inv_dz', T.mc_fp_t),
                        ('n_r', T.mc_size_t), ('n_fi', T.mc_size_t), ('n_z', T.mc_size_t),
                        ('offset', T.mc_size_t), ('k', T.mc_int_t)]
        return ClFluenceCyl

    def __init__(self, raxis=None, fiaxis: Axis = None, zaxis: Axis = None,
                 center: Tuple[float, float] = (0.0, 0.0), mode: str = 'deposition'):
        super().__init__()
        data, nphotons, k = None, 0, _K_DEFAULT
        if isinstance(raxis, FluenceCyl):
            f = raxis
            raxis, fiaxis, zaxis, center = Axis(f.raxis), Axis(f.fiaxis), Axis(f.zaxis), f.center
            nphotons, mode, k = f.nphotons, f.mode, f.k
            if f.raw is not None:
                data = np.copy(f.raw)
        raxis = Axis(0.0, 1.0, 1) if raxis is None else raxis
        fiaxis = Axis(0.0, 2*np.pi, 1) if fiaxis is None else fiaxis
        zaxis = Axis(0.0, 1.0, 1) if zaxis is None else zaxis
        if raxis.logscale or fiaxis.logscale or zaxis.logscale:
            raise ValueError('FluenceCyl does not support logarithmic axes!')
        self._r_axis, self._fi_axis, self._z_axis = raxis, fiaxis, zaxis
        self._center = np.zeros((2,))
        self._center[:] = center
        self._init_common(mode, k, data, nphotons)

    shape = property(lambda self: (self._z_axis.n, self._fi_axis.n, self._r_axis.n))
    raxis = property(lambda self: self._r_axis)
    fiaxis = property(lambda self: self._fi_axis)
    zaxis = property(lambda self: self._z_axis)
    r = property(lambda self: self._r_axis.centers)
    fi = property(lambda self: self._fi_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    center = property(lambda self: self._center)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.center.x, target.center.y = self._center
        target.r_min, target.fi_min, target.z_min = \
            self._r_axis.start, self._fi_axis.start, self._z_axis.start
        target.inv_dr = 1.0/self._r_axis.step
        target.inv_dfi = 1.0/self._fi_axis.step
        target.inv_dz = 1.0/self._z_axis.step
        target.n_r, target.n_fi, target.n_z = self._r_axis.n, self._fi_axis.n, self._z_axis.n
        target.k = self._k
        return target

    def todict(self):
        return {'type': 'FluenceCyl', 'mode': self._mode, 'raxis': self._r_axis.todict(),
                'fiaxis': self._fi_axis.todict(), 'zaxis': self._z_axis.todict(),
                'center': self._center.tolist()}


class FluenceCylt(FluenceCyl):
    """Time-resolved cylindrical grid; raw shape (n_z, n_fi, n_r, n_t)
    (mcfluence/fluencecylt.py).  Time = optical path length / c."""
    cu_type = 'xo::FluCylt'
    _extra_options = [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluenceCylt(cltypes.Structure):
            _fields_ = [('center', T.mc_point2f_t), ('r_min', T.mc_fp_t),
                        ('fi_min', T.mc_fp_t), ('z_min', T.mc_fp_t), ('t_min', T.mc_fp_t),
                        ('inv_dr', T.mc_fp_t), ('inv_dfi', T.mc_fp_t), ('inv_dz', T.mc_fp_t),
                        ('inv_dt', T.mc_fp_t),
                        ('n_r', T.mc_size_t), ('n_fi', T.mc_size_t), ('n_z', T.mc_size_t),
                        ('n_t', T.mc_size_t), ('offset', T.mc_size_t), ('k', T.mc_int_t)]
        return ClFluenceCylt

    def __init__(self, raxis=None, fiaxis: Axis = None, zaxis: Axis = None,
                 taxis: Axis = None, center: Tuple[float, float] = (0.0, 0.0),
                 mode: str = 'deposition'):
        if isinstance(raxis, FluenceCylt):
            taxis = Axis(raxis.taxis)
        taxis = Axis(0.0, 1.0, 1) if taxis is None else taxis
        if taxis.logscale:
            raise ValueError('FluenceCylt does not support logarithmic axes!')
        self._t_axis = taxis
        super().__init__(raxis, fiaxis, zaxis, center, mode)

    shape = property(lambda self: (self._z_axis.n, self._fi_axis.n, self._r_axis.n,
                                   self._t_axis.n))
    taxis = property(lambda self: self._t_axis)
    t = property(lambda self: self._t_axis.centers)
    dt = property(lambda self: abs(self._t_axis.step))

    @property
    def data(self):
        r = self._r_axis.edges
        k = 1.0/(self.nphotons*(r[1:]**2 - r[:-1]**2)*abs(self._fi_axis.step) *
                 abs(self._z_axis.step)*self.dt)
        k.shape = (1, 1, k.size, 1)
        return self._data*k

    def cl_pack(self, mc, target=None):
        target = super().cl_pack(mc, target)
        target.t_min = self._t_axis.start
        target.inv_dt = 1.0/self._t_axis.step
        target.n_t = self._t_axis.n
        return target

    def todict(self):
        d = super().todict()
        d.update(type='FluenceCylt', taxis=self._t_axis.todict())
        return d


class Fluencet(_FluenceBase):
    """Time-resolved x-y-z-t grid; raw shape (nz, ny, nx, nt) (mcfluence/fluencet.py).
    Time = optical path length / c, so the kernel tracks the optical path length."""
    cu_type = 'xo::FluXyzt'
    _extra_options = [('MC_TRACK_OPTICAL_PATHLENGTH', True)]

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClFluencet(cltypes.Structure):
            _fields_ = [('inv_step', T.mc_point4f_t), ('top_left', T.mc_point4f_t),
                        ('shape', T.mc_point4s_t), ('offset', T.mc_size_t),
                        ('k', T.mc_int_t)]
        return ClFluencet

    def __init__(self, xaxis=None, yaxis: Axis = None, zaxis: Axis = None,
                 taxis: Axis = None, mode: str = 'deposition'):
        super().__init__()
        data, nphotons, k = None, 0, _K_DEFAULT
        if isinstance(xaxis, Fluencet):
            f = xaxis
            xaxis, yaxis, zaxis, taxis = (Axis(f.xaxis), Axis(f.yaxis),
                                          Axis(f.zaxis), Axis(f.taxis))
            nphotons, mode, k = f.nphotons, f.mode, f.k
            if f.raw is not None:
                data = np.copy(f.raw)
        xaxis = Axis(-0.5, 0.5, 1) if xaxis is None else xaxis
        yaxis = Axis(-0.5, 0.5, 1) if yaxis is None else yaxis
        zaxis = Axis(0.0, 1.0, 1) if zaxis is None else zaxis
        taxis = Axis(0.0, 1.0, 1) if taxis is None else taxis
        self._x_axis, self._y_axis, self._z_axis, self._t_axis = xaxis, yaxis, zaxis, taxis
        self._init_common(mode, k, data, nphotons)

    shape = property(lambda self: (self._z_axis.n, self._y_axis.n,
                                   self._x_axis.n, self._t_axis.n))
    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    zaxis = property(lambda self: self._z_axis)
    taxis = property(lambda self: self._t_axis)
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    t = property(lambda self: self._t_axis.centers)
    dx = property(lambda self: abs(self._x_axis.step))
    dy = property(lambda self: abs(self._y_axis.step))
    dz = property(lambda self: abs(self._z_axis.step))
    dt = property(lambda self: abs(self._t_axis.step))

    @property
    def data(self):
        return self._data*(1.0/(max(self.nphotons, 1)*self.dx*self.dy*self.dz*self.dt))

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        for c, ax in zip('xyzw', (self._x_axis, self._y_axis, self._z_axis, self._t_axis)):
            setattr(target.top_left, c, ax.start)
            setattr(target.inv_step, c, 1.0/ax.step)
            setattr(target.shape, c, ax.n)
        target.k = self._k
        return target

    def todict(self):
        return {'type': 'Fluencet', 'mode': self._mode, 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'zaxis': self._z_axis.todict(),
                'taxis': self._t_axis.todict()}
