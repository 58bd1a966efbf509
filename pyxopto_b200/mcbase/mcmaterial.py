"""Voxel materials (mirror of ``xopto/mcbase/mcmaterial.py:37-277``)."""
from ..cl import cltypes
from .mcobject import McObject


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


class Materials(McObject):
    def __init__(self, materials):
        super().__init__()
        if isinstance(materials, Materials):
            materials = list(materials)
        self._materials = list(materials)
        self._pf_type = type(self._materials[0].pf)
        for m in self._materials:
            if type(m.pf) is not self._pf_type:
                raise ValueError('All materials must use the same scattering '
                                 'phase function type!')

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
