"""Layer stack of the planar simulator (mirror of ``xopto/mcml/mclayer/layer.py``)."""
from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil import boundary
from ..mcbase.mcmaterial import optical_tensor as _tensor


class Layer(McObject):
    cu_type = 'xo::MlLayer'

    @staticmethod
    def layer_type(mc, pf_type):
        T = mc.types
        class ClLayer(cltypes.Structure):
            _fields_ = [
                ('thickness', T.mc_fp_t), ('top', T.mc_fp_t), ('bottom', T.mc_fp_t),
                ('n', T.mc_fp_t), ('cos_critical_top', T.mc_fp_t),
                ('cos_critical_bottom', T.mc_fp_t), ('mus', T.mc_fp_t),
                ('mua', T.mc_fp_t), ('inv_mut', T.mc_fp_t),
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
        """Packs the fields that do not depend on neighbouring layers
        (layer.py:304-354)."""
        if target is None:
            target = self.fetch_cl_type(mc)()
        mut = self.mua + self.mus
        inv_mut = 1.0/mut if mut > 0.0 else float('inf')
        mua_inv_mut = 1.0 if self.mus == 0.0 else self.mua*inv_mut
        target.thickness, target.n = self.d, self.n
        target.mua, target.mus = self.mua, self.mus
        target.inv_mut, target.mua_inv_mut = inv_mut, mua_inv_mut
        self.pf.cl_pack(mc, target.pf)
        return target

    def todict(self):
        return {'d': self.d, 'n': self.n, 'mua': self.mua, 'mus': self.mus,
                'pf': self.pf.todict(), 'type': type(self).__name__}

    def __repr__(self):
        return 'Layer(d={}, n={}, mua={}, mus={}, pf={})'.format(
            self.d, self.n, self.mua, self.mus, self.pf)


class AnisotropicLayer(McObject):
    """Layer with direction-dependent absorption / scattering coefficients: the
    kernel projects the 3 x 3 tensors on the propagation direction, ``mu(dir) =
    dir^T T dir`` (layer.py:391-790, mcbase.template.h:2227-2230)."""
    cu_type = 'xo::MlAnisoLayer'

    @staticmethod
    def layer_type(mc, pf_type):
        T = mc.types
        class ClAnisotropicLayer(cltypes.Structure):
            _fields_ = [
                ('thickness', T.mc_fp_t), ('top', T.mc_fp_t), ('bottom', T.mc_fp_t),
                ('n', T.mc_fp_t), ('cos_critical_top', T.mc_fp_t),
                ('cos_critical_bottom', T.mc_fp_t), ('mus', T.mc_matrix3f_t),
                ('mua', T.mc_matrix3f_t), ('mut', T.mc_matrix3f_t), ('pf', pf_type)]
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
        target.thickness, target.n = self.d, self.n
        target.mua.fromarray(self._mua)
        target.mus.fromarray(self._mus)
        target.mut.fromarray(self._mua + self._mus)
        self.pf.cl_pack(mc, target.pf)
        return target

    def todict(self):
        return {'d': self.d, 'n': self.n, 'mua': self._mua, 'mus': self._mus,
                'pf': self.pf.todict(), 'type': type(self).__name__}

    @classmethod
    def fromdict(cls, data: dict):
        from ..mcbase import mcpf
        data = dict(data)
        if data.pop('type') != 'AnisotropicLayer':
            raise ValueError('Cannot create an AnisotropicLayer instance from the data!')
        pf_data = data.pop('pf')
        if not hasattr(mcpf, pf_data['type']):
            raise TypeError('Scattering phase function "{}" not implemented'.format(
                pf_data['type']))
        return cls(pf=getattr(mcpf, pf_data['type']).fromdict(pf_data), **data)

    def __repr__(self):
        return 'AnisotropicLayer(d={}, n={}, mua={}, mus={}, pf={})'.format(
            self.d, self.n, self._mua, self._mus, self.pf)


class Layers(McObject):
    """Stack of layers; the first and last describe the surrounding medium."""

    def __init__(self, layers):
        super().__init__()
        if isinstance(layers, Layers):
            layers = list(layers)
        self._layers = list(layers)
        if len(self._layers) < 3:
            raise ValueError('A layer stack needs at least 3 layers '
                             '(2 surrounding + 1 sample layer)!')
        pf_type = type(self._layers[0].pf)
        if any(type(l.pf) is not pf_type for l in self._layers):
            raise ValueError('All the layers must use the same scattering '
                             'phase function model!')
        layer_type = type(self._layers[0])
        for l in self._layers:
            if not isinstance(l, (Layer, AnisotropicLayer)):
                raise TypeError('All the sample layers must be instances of Layer or '
                                'AnisotropicLayer but found {:s}!'.format(type(l).__name__))
            if type(l) is not layer_type:
                raise TypeError('All the sample layers must use the same type!'
                                'Found {} and {}!'.format(layer_type.__name__,
                                                          type(l).__name__))

    def cl_type(self, mc):
        return self._layers[0].fetch_cl_type(mc)*len(self._layers)

    def cl_pack(self, mc, target=None):
        """Packs the stack; top/bottom are accumulated in fp32 exactly like the
        reference (layer.py:981-1000): bottom_i = fp32(fp32(bottom_{i-1}) + d_i)."""
        n_layers = len(self._layers)
        if target is None or len(target) != n_layers:
            target = self.fetch_cl_type(mc)()
        inf = float('inf')
        for i, layer in enumerate(self._layers):
            layer.cl_pack(mc, target[i])
            cc_top = boundary.cos_critical(layer.n, self._layers[i - 1].n) if i > 0 else 0.0
            cc_bottom = boundary.cos_critical(layer.n, self._layers[i + 1].n) \
                if i + 1 < n_layers else 0.0
            if i == 0:
                target[i].top, target[i].bottom, target[i].thickness = -inf, 0.0, inf
            elif i == n_layers - 1:
                target[i].top, target[i].bottom = target[i - 1].bottom, inf
                target[i].thickness = inf
            else:
                target[i].top = target[i - 1].bottom
                target[i].bottom = target[i - 1].bottom + layer.d
            target[i].cos_critical_top = cc_top
            target[i].cos_critical_bottom = cc_bottom
            layer.pf.cl_pack(mc, target[i].pf)
        return target

    def thickness(self) -> float:
        return sum(l.d for l in self._layers[1:-1])

    def layer_index(self, z: float) -> int:
        """Index of the layer that contains depth z."""
        if z < 0.0:
            return 0
        bottom = 0.0
        for i, layer in enumerate(self._layers[1:-1], start=1):
            bottom += layer.d
            if z < bottom:
                return i
        return len(self._layers) - 1

    def layer(self, index):
        return self._layers[index]

    def __getitem__(self, i):
        return self._layers[i]

    def __len__(self):
        return len(self._layers)

    def __iter__(self):
        return iter(self._layers)

    def todict(self):
        return {'layers': [l.todict() for l in self._layers], 'type': 'Layers'}
