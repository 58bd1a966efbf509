"""Voxel materials (mirror of ``xopto/mcbase/mcmaterial.py:37-277`` and the
anisotropic variant, ``:310-655``)."""
import numpy as np

from ..cl import cltypes
from .mcobject import McObject


def optical_tensor(value) -> np.ndarray:
    """3 x 3 coefficient tensor from a scalar (isotropic medium), the 3 diagonal
    elements, or the full matrix (mcmaterial.py:545-590, mclayer/layer.py:662-708)."""
    t = np.zeros((3, 3))
    if isinstance(value, (float, int)):
        t[0, 0] = t[1, 1] = t[2, 2] = value
    else:
        value = np.asarray(value, dtype=float)
        if value.size == 3:
            t[0, 0], t[1, 1], t[2, 2] = value.ravel()
        else:
            t[:] = value
    return t


class Material(McObject):
    @staticmethod
    def material_type(mc, pf_type):
        T = mc.types
        class ClMaterial(cltypes.Structure):
            _fields_ = [('n', T.mc_fp_t), ('mus', T.mc_fp_t), ('mua', T.mc_fp_t),
                        ('inv_mut', T.mc_fp_t), ('mua_inv_mut', T.mc_fp_t),
                        ('pf', pf_type)]
        return ClMaterial

    def cl_type(self, mc):
        return self.material_type(mc, self.pf.fetch_cl_type(mc))

    def __init__(self, n: float, mua: float, mus: float, pf):
        super().__init__()
        self.n, self.mua, self.mus = float(n), float(mua), float(mus)
        self._pf = pf

    def _set_pf(self, pf):
        if type(self._pf) is not type(pf):
            raise ValueError('The scattering phase function type of the '
                             'material must not change!')
        self._pf = pf

    pf = property(lambda self: self._pf, _set_pf)

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.fetch_cl_type(mc)()
        mut = self.mua + self.mus
        inv_mut = 1.0/mut if mut > 0.0 else float('inf')
        mua_inv_mut = 1.0 if self.mus == 0.0 else self.mua*inv_mut
        target.n, target.mua, target.mus = self.n, self.mua, self.mus
        target.inv_mut, target.mua_inv_mut = inv_mut, mua_inv_mut
        self.pf.cl_pack(mc, target.pf)
        return target

    def todict(self):
        return {'n': self.n, 'mua': self.mua, 'mus': self.mus,
                'pf': self.pf.todict(), 'type': type(self).__name__}

    def __repr__(self):
        return 'Material(n={}, mua={}, mus={}, pf={})'.format(
            self.n, self.mua, self.mus, self.pf)


class AnisotropicMaterial(McObject):
    """Material with absorption / scattering tensors; the kernel projects them on the
    propagation direction, ``mu(dir) = dir^T T dir`` (mcmaterial.py:310-655)."""
    @staticmethod
    def material_type(mc, pf_type):
        T = mc.types
        class ClAnisotropicMaterial(cltypes.Structure):
            _fields_ = [('n', T.mc_fp_t), ('mus', T.mc_matrix3f_t),
                        ('mua', T.mc_matrix3f_t), ('mut', T.mc_matrix3f_t),
                        ('pf', pf_type)]
        return ClAnisotropicMaterial

    def cl_type(self, mc):
        return self.material_type(mc, self.pf.fetch_cl_type(mc))

    def __init__(self, n: float, mua, mus, pf):
        super().__init__()
        self.n = float(n)
        self._mua, self._mus = optical_tensor(mua), optical_tensor(mus)
        self._pf = pf

    def _set_mua(self, mua):
        self._mua = optical_tensor(mua)

    def _set_mus(self, mus):
        self._mus = optical_tensor(mus)

    mua = property(lambda self: self._mua, _set_mua, None,
                   'Absorption coefficient tensor (1/m).')
    mus = property(lambda self: self._mus, _set_mus, None,
                   'Scattering coefficient tensor (1/m).')

    def _set_pf(self, pf):
        if type(self._pf) is not type(pf):
            raise ValueError('The scattering phase function type of the '
                             'material must not change!')
        self._pf = pf

    pf = property(lambda self: self._pf, _set_pf)

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
        return {'n': self.n, 'mua': self._mua.tolist(), 'mus': self._mus.tolist(),
                'pf': self.pf.todict(), 'type': type(self).__name__}

    def __repr__(self):
        return 'AnisotropicMaterial(n={}, mua={}, mus={}, pf={})'.format(
            self.n, self._mua, self._mus, self.pf)


class Materials(McObject):
    def __init__(self, materials):
        super().__init__()
        if isinstance(materials, Materials):
            materials = list(materials)
        self._materials = list(materials)
        self._pf_type = type(self._materials[0].pf)
        material_type = type(self._materials[0])
        for m in self._materials:
            if type(m.pf) is not self._pf_type:
                raise ValueError('All materials must use the same scattering '
                                 'phase function type!')
            if type(m) is not material_type:
                raise TypeError('All materials must be of the same type!')

    def cl_type(self, mc):
        return self._materials[0].fetch_cl_type(mc)*len(self._materials)

    def cl_pack(self, mc, target=None):
        if target is None or len(target) != len(self._materials):
            target = self.fetch_cl_type(mc)()
        for i, m in enumerate(self._materials):
            m.cl_pack(mc, target[i])
        return target

    def material(self, index):
        return self._materials[index]

    def __getitem__(self, i):
        return self._materials[i]

    def __len__(self):
        return len(self._materials)

    def __iter__(self):
        return iter(self._materials)

    def todict(self):
        return {'materials': [m.todict() for m in self._materials], 'type': 'Materials'}
