"""Concentric layer stack of the cylindrical simulator (mirror of
``xopto/mccyl/mclayer/layer.py``).  The first layer is the surrounding medium,
the following layers are ordered from the outermost to the innermost; ``d`` is a
layer *diameter*."""
from typing import Tuple

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil import boundary
from ..mcbase.mcmaterial import optical_tensor as _tensor


def ray_cylinder_intersection(r: float, pos, dir) -> Tuple[float, float]:
    """Distances to the two intersections of a ray with the cylinder
    x^2 + y^2 = r^2, or (None, None) (layer.py:36-94)."""
    a = dir[0]**2 + dir[1]**2
    b = 2.0*(pos[0]*dir[0] + pos[1]*dir[1])
    c = pos[0]**2 + pos[1]**2 - r**2
    D = b**2 - 4*a*c
    if D < 0.0 or a == 0.0:
        return None, None
    D = D**0.5
    return (-b - D)/(2.0*a), (-b + D)/(2.0*a)


class Layer(McObject):
    cu_type = 'xo::CylLayer'

    @staticmethod
    def layer_type(mc, pf_type):
        T = mc.types
        class ClLayer(cltypes.Structure):
            _fields_ = [
                ('r_inner', T.mc_fp_t), ('r_outer', T.mc_fp_t), ('n', T.mc_fp_t),
                ('cos_critical_inner', T.mc_fp_t), ('cos_critical_outer', T.mc_fp_t),
                ('mus', T.mc_fp_t), ('mua', T.mc_fp_t), ('inv_mut', T.mc_fp_t),
                ('mua_inv_mut', T.mc_fp_t), ('pf', pf_type)]
        return ClLayer

    def cl_type(self, mc):
        return self.layer_type(mc, self.pf.fetch_cl_type(mc))

    def __init__(self, d: float, n: float, mua: float, mus: float, pf):
        super().__init__()
        self.d, self.n, self.mua, self.mus = float(d), float(n), float(mua), float(mus)
        self._pf = pf

    def _set_pf(self, pf):
        if type(self._pf) is not type(pf):
            raise ValueError('The scattering phase function type '
                             'of the layer must not change!')
        self._pf = pf

    pf = property(lambda self: self._pf, _set_pf, None, 'Phase function object.')

    def cl_pack(self, mc, target=None):
        """Fields that do not depend on the neighbours (layer.py:369-424)."""
        if target is None:
            target = self.fetch_cl_type(mc)()
        mut = self.mua + self.mus
        inv_mut = 1.0/mut if mut > 0.0 else float('inf')
        mua_inv_mut = 1.0 if self.mus == 0.0 else self.mua*inv_mut
        target.n = self.n
        target.mua, target.mus = self.mua, self.mus
        target.inv_mut, target.mua_inv_mut = inv_mut, mua_inv_mut
        self.pf.cl_pack(mc, target.pf)
        return target

    def todict(self):
        return {'d': self.d, 'n': self.n, 'mua': self.mua, 'mus': self.mus,
                'pf': self.pf.todict(), 'type': 'Layer'}

    def __repr__(self):
        return 'Layer(d={}, n={}, mua={}, mus={}, pf={})'.format(
            self.d, self.n, self.mua, self.mus, self.pf)


class AnisotropicLayer(McObject):
    """Concentric layer with absorption / scattering tensors projected on the
    propagation direction (layer.py:455-760)."""
    cu_type = 'xo::CylLayer'

    @staticmethod
    def layer_type(mc, pf_type):
        T = mc.types
        class ClAnisotropicLayer(cltypes.Structure):
            _fields_ = [
                ('r_inner', T.mc_fp_t), ('r_outer', T.mc_fp_t), ('n', T.mc_fp_t),
                ('cos_critical_inner', T.mc_fp_t), ('cos_critical_outer', T.mc_fp_t),
                ('mus', T.mc_matrix3f_t), ('mua', T.mc_matrix3f_t),
                ('mut', T.mc_matrix3f_t), ('pf', pf_type)]
        return ClAnisotropicLayer

    def cl_type(self, mc):
        return self.layer_type(mc, self.pf.fetch_cl_type(mc))

    def __init__(self, d: float, n: float, mua, mus, pf):
        super().__init__()
        self.d, self.n = float(d), float(n)
        self._mua, self._mus = _tensor(mua), _tensor(mus)
        self._pf = pf

    def _set_mua(self, mua):
        self._mua = _tensor(mua)

    def _set_mus(self, mus):
        self._mus = _tensor(mus)

    mua = property(lambda self: self._mua, _set_mua, None,
                   'Absorption coefficient tensor (3x3) of the layer (1/m).')
    mus = property(lambda self: self._mus, _set_mus, None,
                   'Scattering coefficient tensor (3x3) of the layer (1/m).')

    def _set_pf(self, pf):
        if type(self._pf) is not type(pf):
            raise ValueError('The scattering phase function type '
                             'of the layer must not change!')
        self._pf = pf

    pf = property(lambda self: self._pf, _set_pf, None, 'Phase function object.')

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        target.n = self.n
        target.mua.fromarray(self._mua)
        target.mus.fromarray(self._mus)
        target.mut.fromarray(self._mua + self._mus)
        self.pf.cl_pack(mc, target.pf)
        return target

    def todict(self):
        return {'d': self.d, 'n': self.n, 'mua': self._mua.tolist(), 'mus': self._mus.tolist(),
                'pf': self.pf.todict(), 'type': 'AnisotropicLayer'}

    def __repr__(self):
        return 'AnisotropicLayer(d={}, n={}, mua={}, mus={}, pf={})'.format(
            self.d, self.n, self._mua, self._mus, self.pf)


class Layers(McObject):
    def __init__(self, layers):
        super().__init__()
        if isinstance(layers, Layers):
            layers = layers.tolist()
        self._layers = list(layers)
        self.check()

    def check(self):
        if len(self._layers) < 2:
            raise ValueError('At least two layers are required, '
                             'but got only {:d}!'.format(len(self._layers)))
        pf_type = type(self._layers[1].pf)
        layer_type = type(self._layers[0])
        for layer in self._layers:
            if not isinstance(layer, (Layer, AnisotropicLayer)):
                raise TypeError('All layers must be instances of Layer or AnisotropicLayer '
                                'but found {:s}!'.format(type(layer).__name__))
            if type(layer) is not layer_type:
                raise TypeError('All the sample layers must use the same type!')
            if type(layer.pf) is not pf_type:
                raise TypeError('All the sample layer must use the same scattering '
                                'phase function model!')
        d_prev = float('inf')
        for layer in self._layers[1:]:
            if layer.d > d_prev:
                raise ValueError('The diameters of layers must be '
                                 'monotonically decreasing!')
            d_prev = layer.d

    def layer(self, index: int) -> Layer:
        return self._layers[index]

    def layer_index(self, r: float) -> int:
        """Index of the layer that contains radius r; [r_inner, r_outer)
        (layer.py:951-975)."""
        index = 0
        for pos, layer in enumerate(self._layers[::-1]):
            if 2.0*r < layer.d:
                index = len(self._layers) - 1 - pos
                break
        return index

    def diameter(self) -> float:
        return self._layers[1].d

    def cl_type(self, mc):
        return self._layers[0].fetch_cl_type(mc)*len(self._layers)

    def cl_pack(self, mc, target=None):
        """Radii and critical cosines from the neighbours (layer.py:1037-1092)."""
        self.check()
        n_layers = len(self._layers)
        if target is None or len(target) != n_layers:
            target = self.fetch_cl_type(mc)()
        for i, layer in enumerate(self._layers):
            layer.cl_pack(mc, target[i])
            cc_outer = cc_inner = 0.0
            if i > 0:
                cc_outer = boundary.cos_critical(layer.n, self._layers[i - 1].n)
            if i + 1 < n_layers:
                cc_inner = boundary.cos_critical(layer.n, self._layers[i + 1].n)
            if i == 0:
                target[i].r_outer = float('inf')
                target[i].r_inner = self._layers[1].d*0.5
            else:
                target[i].r_outer = layer.d*0.5
                target[i].r_inner = self._layers[i + 1].d*0.5 if i + 1 < n_layers else 0.0
            target[i].cos_critical_outer = cc_outer
            target[i].cos_critical_inner = cc_inner
            layer.pf.cl_pack(mc, target[i].pf)
        return target

    def intersect(self, pos, dir, entrance: bool = False):
        """Intersection of a ray with the sample surface and the (unnormalised,
        inward) surface normal there (layer.py:1094-1143)."""
        d1, d2 = ray_cylinder_intersection(self.layer(1).d*0.5, pos, dir)
        if d1 is None or d2 is None:
            return None, None
        if d1 < 0.0 and d2 < 0.0 and not entrance:
            return None, None
        if entrance:
            d = min(d1, d2)
        elif d1 > 0.0 and d2 >= 0.0:
            d = min(d1, d2)
        else:
            d = max(d1, d2)
        intersection = pos[0] + dir[0]*d, pos[1] + dir[1]*d, pos[2] + dir[2]*d
        normal = (-intersection[0], -intersection[1], 0.0)
        return intersection, normal

    def tolist(self):
        return list(self._layers)

    def todict(self):
        return {'layers': [l.todict() for l in self._layers], 'type': 'Layers'}

    def __getitem__(self, i):
        return self._layers[i]

    def __setitem__(self, i, v):
        self._layers[i] = v

    def __len__(self):
        return len(self._layers)

    def __iter__(self):
        return iter(self._layers)
