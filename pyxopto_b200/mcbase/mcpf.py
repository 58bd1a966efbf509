"""Scattering phase functions (mirror of ``xopto/mcbase/mcpf``: Hg, MHg, Hg2, Gk,
MGk, Gk2, Pc, MPc, Lut, LutEx).

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


class HgDir(PfBase):
    """Henyey-Greenstein with a preferred scattering direction (mcpf/hgdir.py): with
    probability ``p`` the deflection is measured from ``direction`` instead of the
    packet's direction.  Samples the new direction itself
    (``MC_PF_SAMPLE_DIRECTION``)."""
    cu_type = 'xo::PfHgDir'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClHgDir(cltypes.Structure):
            _fields_ = [('direction', T.mc_point3f_t), ('g', T.mc_fp_t), ('p', T.mc_fp_t)]
        return ClHgDir

    @staticmethod
    def cl_options(mc):
        return [('MC_PF_SAMPLE_DIRECTION', True)]

    def __init__(self, g: float, direction=(0.0, 0.0, 1.0), p: float = 1.0):
        super().__init__()
        self._direction = np.array((0.0, 0.0, 1.0))
        self.g, self.p, self.direction = g, p, direction

    def _set_g(self, g):
        self._g = min(max(float(g), -1.0), 1.0)

    def _set_p(self, p):
        self._p = min(max(float(p), 0.0), 1.0)

    def _set_direction(self, d):
        self._direction[:] = d
        norm = np.linalg.norm(self._direction)
        if norm == 0.0:
            raise ValueError('Direction vector norm/length must not be 0!')
        self._direction *= 1.0/norm

    g = property(lambda self: self._g, _set_g)
    p = property(lambda self: self._p, _set_p)
    direction = property(lambda self: self._direction, _set_direction)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.g = self._g
        target.p = self._p
        target.direction.fromarray(self._direction)
        return target

    def todict(self):
        return {'g': self._g, 'p': self._p, 'direction': self._direction.tolist(),
                'type': type(self).__name__}


class MHg(PfBase):
    """Modified Henyey-Greenstein (mcpf/mhg.py)."""
    cu_type = 'xo::PfMHg'

    @staticmethod
    def cl_type(mc):
        class ClMHg(cltypes.Structure):
            _fields_ = [('g', mc.types.mc_fp_t), ('beta', mc.types.mc_fp_t)]
        return ClMHg

    def __init__(self, g: float, b: float = None, beta: float = None):
        # (the reference names the Rayleigh fraction ``b``, mhg.py:112; ``beta`` is the
        # name of the packed field)
        super().__init__()
        if b is None:
            b = beta
        if b is None:
            raise TypeError("MHg() missing 1 required positional argument: 'b'")
        self.g = g
        self.beta = b

    def _set_g(self, g):
        self._g = min(max(float(g), -1.0), 1.0)

    def _set_beta(self, b):
        self._beta = min(max(float(b), 0.0), 1.0)

    g = property(lambda self: self._g, _set_g)
    beta = property(lambda self: self._beta, _set_beta)
    b = property(lambda self: self._beta, _set_beta)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.g = self._g
        target.beta = self._beta
        return target

    def todict(self):
        return {'g': self._g, 'b': self._beta, 'type': type(self).__name__}

    def __repr__(self):
        return 'MHg(g={}, b={})'.format(self._g, self._beta)


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


class Hg2(PfBase):
    """Two-term Henyey-Greenstein (mcpf/hg2.py): (1 - b) Hg(g1 >= 0) + b Hg(g2 <= 0)."""
    cu_type = 'xo::PfHg2'

    @staticmethod
    def cl_type(mc):
        T_HG = Hg.cl_type(mc)
        class ClHg2(cltypes.Structure):
            _fields_ = [('hg_1', T_HG), ('hg_2', T_HG), ('b', mc.types.mc_fp_t)]
        return ClHg2

    def __init__(self, g1: float, g2: float, b: float):
        super().__init__()
        self._hg1, self._hg2 = Hg(g1), Hg(g2)
        self.g1, self.g2, self.b = g1, g2, b

    def _set_g1(self, g):
        self._hg1.g = min(max(float(g), 0.0), 1.0)

    def _set_g2(self, g):
        self._hg2.g = min(max(float(g), -1.0), 0.0)

    def _set_b(self, b):
        self._b = min(max(float(b), 0.0), 1.0)

    g1 = property(lambda self: self._hg1.g, _set_g1, None, 'Anisotropy of the first HG.')
    g2 = property(lambda self: self._hg2.g, _set_g2, None, 'Anisotropy of the second HG.')
    b = property(lambda self: self._b, _set_b, None,
                 'Relative contribution of the second HG term.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        self._hg1.cl_pack(mc, target.hg_1)
        self._hg2.cl_pack(mc, target.hg_2)
        target.b = self._b
        return target

    def todict(self):
        return {'g1': self.g1, 'g2': self.g2, 'b': self._b, 'type': type(self).__name__}

    def __repr__(self):
        return 'Hg2(g1={}, g2={}, b={})'.format(self.g1, self.g2, self._b)


class Gk2(PfBase):
    """Two-term Gegenbauer kernel (mcpf/gk2.py)."""
    cu_type = 'xo::PfGk2'

    @staticmethod
    def cl_type(mc):
        T_GK = Gk.cl_type(mc)
        class ClGk2(cltypes.Structure):
            _fields_ = [('gk_1', T_GK), ('gk_2', T_GK), ('b', mc.types.mc_fp_t)]
        return ClGk2

    def __init__(self, g1: float, a1: float, g2: float, a2: float, b: float):
        super().__init__()
        self._gk1, self._gk2 = Gk(g1, a1), Gk(g2, a2)
        self.g1, self.g2, self.a1, self.a2, self.b = g1, g2, a1, a2, b

    def _set_g1(self, g):
        self._gk1.g = min(max(float(g), 0.0), 1.0)

    def _set_a1(self, a):
        self._gk1.a = max(float(a), -0.5)

    def _set_g2(self, g):
        self._gk2.g = min(max(float(g), -1.0), 0.0)

    def _set_a2(self, a):
        self._gk2.a = max(float(a), -0.5)

    def _set_b(self, b):
        self._b = min(max(float(b), 0.0), 1.0)

    g1 = property(lambda self: self._gk1.g, _set_g1)
    a1 = property(lambda self: self._gk1.a, _set_a1)
    g2 = property(lambda self: self._gk2.g, _set_g2)
    a2 = property(lambda self: self._gk2.a, _set_a2)
    b = property(lambda self: self._b, _set_b, None,
                 'Relative contribution of the second GK term.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        self._gk1.cl_pack(mc, target.gk_1)
        self._gk2.cl_pack(mc, target.gk_2)
        target.b = self._b
        return target

    def todict(self):
        return {'g1': self.g1, 'a1': self.a1, 'g2': self.g2, 'a2': self.a2, 'b': self._b,
                'type': type(self).__name__}

    def __repr__(self):
        return 'Gk2(g1={}, a1={}, g2={}, a2={}, b={})'.format(
            self.g1, self.a1, self.g2, self.a2, self._b)


class MGk(Gk):
    """Modified Gegenbauer kernel (mcpf/mgk.py): b Gk(g, a) + (1 - b) 3/2 cos^2."""
    cu_type = 'xo::PfMGk'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClMGk(cltypes.Structure):
            _fields_ = [('g', T.mc_fp_t), ('a', T.mc_fp_t), ('beta', T.mc_fp_t),
                        ('inv_a', T.mc_fp_t), ('a1', T.mc_fp_t), ('a2', T.mc_fp_t)]
        return ClMGk

    def __init__(self, g: float, a: float, b: float):
        super().__init__(g, a)
        self.b = b

    def _set_b(self, b):
        self._b = min(max(float(b), 0.0), 1.0)

    b = property(lambda self: self._b, _set_b, None,
                 'Contribution of the Gegenbauer kernel term.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        inv_a, a1, a2 = self._precalculated()
        target.g, target.a, target.beta = self._g, self._a, self._b
        target.inv_a, target.a1, target.a2 = inv_a, a1, a2
        return target

    def todict(self):
        return {'g': self._g, 'a': self._a, 'b': self._b, 'type': type(self).__name__}

    def __repr__(self):
        return 'MGk(g={}, a={}, b={})'.format(self._g, self._a, self._b)


class Pc(PfBase):
    """Power of cosines (mcpf/pc.py)."""
    cu_type = 'xo::PfPc'

    @staticmethod
    def cl_type(mc):
        class ClPc(cltypes.Structure):
            _fields_ = [('n', mc.types.mc_fp_t)]
        return ClPc

    def __init__(self, n: float):
        super().__init__()
        self.n = n

    def _set_n(self, n):
        self._n = float(n)

    n = property(lambda self: self._n, _set_n, None, 'Power of cosine.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.n = self._n
        return target

    def todict(self):
        return {'n': self._n, 'type': type(self).__name__}

    def __repr__(self):
        return 'Pc(n={})'.format(self._n)


class Rayleigh(PfBase):
    """Rayleigh scattering (mcpf/rayleigh.py): p ~ (1 + 3 gamma) + (1 - gamma) cos^2.  The
    reference packs it (rayleigh.py:54-58, 166-176) but its sampling text does not compile
    (rayleigh.py:96-100); the kernel here samples it with the evident meaning of that text."""
    cu_type = 'xo::PfRayleigh'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClRayleigh(cltypes.Structure):
            _fields_ = [('gamma', T.mc_fp_t), ('a', T.mc_fp_t), ('b', T.mc_fp_t)]
        return ClRayleigh

    def __init__(self, gamma: float):
        super().__init__()
        self.gamma = gamma

    def _set_gamma(self, gamma):
        self._gamma = min(max(float(gamma), 0.0), 1.0)

    gamma = property(lambda self: self._gamma, _set_gamma, None, 'Molecular anisotropy factor.')

    def _precalculated(self):
        gamma = self._gamma
        if gamma == 1.0:
            return 0.0, 0.0
        return 3.0*(1.0 + 3.0*gamma)/(1.0 - gamma), 4.0*(1.0 + 2.0*gamma)/(1.0 - gamma)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.gamma = self._gamma
        target.a, target.b = self._precalculated()
        return target

    def todict(self):
        return {'gamma': self._gamma, 'type': type(self).__name__}

    def __repr__(self):
        return 'Rayleigh(gamma={})'.format(self._gamma)


class MPc(Pc):
    """Modified power of cosines (mcpf/mpc.py): b Pc(n) + (1 - b) 3/2 cos^2."""
    cu_type = 'xo::PfMPc'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClMPc(cltypes.Structure):
            _fields_ = [('n', T.mc_fp_t), ('beta', T.mc_fp_t)]
        return ClMPc

    def __init__(self, n: float, b: float):
        super().__init__(n)
        self.b = b

    def _set_b(self, b):
        self._b = min(max(float(b), 0.0), 1.0)

    b = property(lambda self: self._b, _set_b, None,
                 'Contribution of the power of cosine term.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.n, target.beta = self._n, self._b
        return target

    def todict(self):
        return {'n': self._n, 'b': self._b, 'type': type(self).__name__}

    def __repr__(self):
        return 'MPc(n={}, b={})'.format(self._n, self._b)


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
