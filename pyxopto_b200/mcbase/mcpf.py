"""Scattering phase functions (mirror of ``xopto/mcbase/mcpf``: Hg, MHg, Gk, Lut, LutEx).

Each class packs the reference's ``McPf`` struct and names the CUDA struct in
``csrc/kernels/xo_pf.cuh`` that samples it.
"""
import numpy as np

from ..cl import cltypes
from .mcobject import McObject


class PfBase(McObject):
    def cl_pack(self, mc, target=None):
        raise NotImplementedError

    def todict(self) -> dict:
        raise NotImplementedError


class Hg(PfBase):
    """Henyey-Greenstein (mcpf/hg.py)."""
    cu_type = 'xo::PfHg'

    @staticmethod
    def cl_type(mc):
        class ClHg(cltypes.Structure):
            _fields_ = [('g', mc.types.mc_fp_t)]
        return ClHg

    def __init__(self, g: float):
        super().__init__()
        self.g = g

    def _set_g(self, g):
        self._g = min(max(float(g), -1.0), 1.0)

    g = property(lambda self: self._g, _set_g, None, 'Anisotropy factor.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.g = self._g
        return target

    def todict(self):
        return {'g': self._g, 'type': type(self).__name__}

    def __repr__(self):
        return 'Hg(g={})'.format(self._g)


class MHg(PfBase):
    """Modified Henyey-Greenstein (mcpf/mhg.py)."""
    cu_type = 'xo::PfMHg'

    @staticmethod
    def cl_type(mc):
        class ClMHg(cltypes.Structure):
            _fields_ = [('g', mc.types.mc_fp_t), ('beta', mc.types.mc_fp_t)]
        return ClMHg

    def __init__(self, g: float, beta: float):
        super().__init__()
        self.g = g
        self.beta = beta

    def _set_g(self, g):
        self._g = min(max(float(g), -1.0), 1.0)

    def _set_beta(self, b):
        self._beta = min(max(float(b), 0.0), 1.0)

    g = property(lambda self: self._g, _set_g)
    beta = property(lambda self: self._beta, _set_beta)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.g = self._g
        target.beta = self._beta
        return target

    def todict(self):
        return {'g': self._g, 'beta': self._beta, 'type': type(self).__name__}

    def __repr__(self):
        return 'MHg(g={}, beta={})'.format(self._g, self._beta)


class Gk(PfBase):
    """Gegenbauer kernel (mcpf/gk.py); host precompute as gk.py:168-192."""
    cu_type = 'xo::PfGk'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClGk(cltypes.Structure):
            _fields_ = [('g', T.mc_fp_t), ('a', T.mc_fp_t), ('inv_a', T.mc_fp_t),
                        ('a1', T.mc_fp_t), ('a2', T.mc_fp_t)]
        return ClGk

    def __init__(self, g: float, a: float):
        super().__init__()
        self._g = min(max(float(g), -1.0), 1.0)
        self._a = max(float(a), -0.5)

    def _set_g(self, g):
        self._g = min(max(float(g), -1.0), 1.0)

    def _set_a(self, a):
        self._a = max(float(a), -0.5)

    g = property(lambda self: self._g, _set_g)
    a = property(lambda self: self._a, _set_a)

    def _precalculated(self):
        g, a = self._g, self._a
        if g == 0:
            return 0.0, 0.0, 0.0
        if a == 0:
            return 0.0, (1 + g**2)/(2*g), (1 + 2*g + g**2)/(2*g)
        temp = a*g*(1.0 - g*g)**(2.0*a)
        temp = temp/(np.pi*((1.0 + g)**(2.0*a) - (1.0 - g)**(2.0*a)))
        return 1.0/a, 2.0*a*g/(2.0*np.pi*temp), (1 + g)**(-2.0*a)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        inv_a, a1, a2 = self._precalculated()
        target.g, target.a = self._g, self._a
        target.inv_a, target.a1, target.a2 = inv_a, a1, a2
        return target

    def todict(self):
        return {'g': self._g, 'a': self._a, 'type': type(self).__name__}

    def __repr__(self):
        return 'Gk(g={}, a={})'.format(self._g, self._a)


class Lut(PfBase):
    """Lookup-table phase function (mcpf/lut.py:33-256): the deflection cosine
    is interpolated at ``index = (a/(xi - c) - b + 1)*(size - 1)/2``."""
    cu_type = 'xo::PfLut'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClLut(cltypes.Structure):
            _fields_ = [('a', T.mc_fp_t), ('b', T.mc_fp_t), ('c', T.mc_fp_t),
                        ('offset', T.mc_size_t), ('size', T.mc_size_t)]
        return ClLut

    @staticmethod
    def cl_options(mc):
        return [('MC_USE_FP_LUT', True)]

    def __init__(self, params, lut: np.ndarray):
        super().__init__()
        self._offset = 0
        self._lut = np.asarray(lut, dtype=np.float64)
        self._params = np.zeros((3,))
        self._params[:] = params

    params = property(lambda self: self._params)
    lut = property(lambda self: self._lut)
    size = property(lambda self: self._lut.size)
    offset = property(lambda self: self._offset)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        entry = mc.append_r_lut(self._lut)
        self._offset = entry.offset
        target.a, target.b, target.c = self._params
        target.offset = self._offset
        target.size = self._lut.size
        return target

    def todict(self):
        return {'params': self._params, 'lut': self._lut, 'type': type(self).__name__}

    def __repr__(self):
        return 'Lut(params={}, size={})'.format(tuple(self._params), self._lut.size)


class LutEx(Lut):
    """Lut built from a host-side phase-function object exposing
    ``mclut(lutsize, **kwargs) -> (params, lut)`` (e.g. ``xopto.pf.Hg`` of the
    reference; the host maths of ``xopto.pf`` is reused, not rebuilt - SURVEY 2.1 #9)."""

    def __init__(self, pftype, pfargs, lutsize: int = 2000, **kwargs):
        if isinstance(pftype, str):
            raise ValueError('LutEx needs a phase-function class with an '
                             'mclut() method (e.g. from xopto.pf), got a name.')
        self._pfargs = tuple(pfargs)
        self._pf_obj = pftype(*pfargs)
        params, lut = self._pf_obj.mclut(lutsize, **kwargs)
        super().__init__(params, lut)

    pfargs = property(lambda self: self._pfargs)
