"""Read-only lookup-table pool (mirror of ``xopto/mcbase/mcutil/lut.py:408-578``):
tables are de-duplicated by value and concatenated into one float array."""
import numpy as np

from ...cl import cltypes
from ..mcobject import McObject


class LutEntry:
    def __init__(self, manager, data: np.ndarray, offset: int):
        self.manager = manager
        self.data = data
        self.offset = int(offset)
        self.size = int(data.size)


class LutManager:
    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.clear()

    def clear(self):
        self._entries = []
        self._size = 0

    def append(self, data: np.ndarray, force: bool = False) -> LutEntry:
        flat = np.asarray(data).ravel()
        if not force:
            for e in self._entries:
                if e.data.shape == flat.shape and np.array_equal(e.data, flat):
                    return e
        entry = LutEntry(self, flat, self._size)
        self._entries.append(entry)
        self._size += flat.size
        return entry

    def pack_into(self, target: np.ndarray = None) -> np.ndarray:
        if target is None or target.size < self._size:
            target = np.zeros(max(self._size, 1), dtype=self.dtype)
        for e in self._entries:
            target[e.offset:e.offset + e.size] = e.data
        return target

    @property
    def size(self) -> int:
        return self._size

    def __len__(self):
        return len(self._entries)


class LinearLut(McObject):
    """Lookup table with uniformly spaced entries between ``first`` and ``last``
    (mcutil/lut.py:33-238); packs the kernel's ``mc_fp_lut_t`` descriptor and puts
    the table into the simulator's float pool."""
    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClLinearLut(cltypes.Structure):
            _fields_ = [('first', T.mc_fp_t), ('inv_span', T.mc_fp_t),
                        ('n', T.mc_size_t), ('offset', T.mc_size_t)]
        return ClLinearLut

    @staticmethod
    def fromfile(filename: str) -> 'LinearLut':
        data = np.load(filename)
        return LinearLut(data['lut_data'], data['first'], data['last'])

    def __init__(self, lut_data, first: float = 0.0, last: float = 1.0):
        if isinstance(lut_data, LinearLut):
            lut = lut_data
            self._first, self._last = lut.first, lut.last
            self._lut_data = np.copy(lut.data)
        elif isinstance(lut_data, str):
            np_data = np.load(lut_data)
            self._first, self._last = float(np_data['first']), float(np_data['last'])
            self._lut_data = np.asarray(np_data['lut_data'], dtype=np.float64)
        else:
            self._first, self._last = float(first), float(last)
            self._lut_data = np.asarray(lut_data, dtype=np.float64)

    first = property(lambda self: self._first)
    last = property(lambda self: self._last)
    span = property(lambda self: self._last - self._first)
    data = property(lambda self: self._lut_data)

    def __call__(self, x):
        n = self._lut_data.size
        fp_ind = (np.asarray(x) - self._first)/(self._last - self._first)*(n - 1)
        ind_1 = np.clip(np.floor(fp_ind), 0, n - 1).astype(np.intp)
        ind_2 = np.clip(ind_1 + 1, 0, n - 1)
        w = fp_ind - ind_1
        return (1.0 - w)*self._lut_data[ind_1] + w*self._lut_data[ind_2]

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        entry = mc.append_r_lut(self._lut_data)
        target.offset = entry.offset
        target.first = self.first
        target.inv_span = 1.0/(self.last - self.first)
        target.n = self.data.size
        return target

    def todict(self, np2list: bool = True) -> dict:
        return {'type': type(self).__name__, 'first': self._first, 'last': self._last,
                'lut_data': self._lut_data.tolist() if np2list else self._lut_data}

    @classmethod
    def fromdict(cls, data: dict):
        data = dict(data)
        data.pop('type', None)
        return LinearLut(**data)

    def save(self, filename: str):
        np.savez_compressed(filename, **self.todict(False))


class CollectionLut(LinearLut):
    """Angular sensitivity of a detector as a function of the incidence-angle
    cosine, resampled to ``n`` uniformly spaced cosines (mcutil/lut.py:327-362)."""
    def __init__(self, sensitivity, costheta=None, n: int = 1000):
        if isinstance(sensitivity, (str, CollectionLut)):
            super().__init__(sensitivity)
            return
        from scipy.interpolate import interp1d
        sensitivity = np.asarray(sensitivity, dtype=np.float64)
        if costheta is None:
            costheta = np.linspace(0.0, 1.0, sensitivity.size)
        else:
            costheta = np.asarray(costheta)
        ct = np.linspace(costheta.min(), costheta.max(), n)
        super().__init__(interp1d(costheta, sensitivity)(ct), ct[0], ct[-1])


class EmissionLut(LinearLut):
    """Table for sampling the emission-angle cosine of a source with the given
    angular radiance (azimuthal symmetry): sampled with a uniform random number
    from [0, 1] (mcutil/lut.py:240-325; CDF by Simpson's rule or adaptive
    quadrature)."""
    def __init__(self, radiance, costheta=None, n: int = 2000, npts: int = 10000,
                 meth: str = 'simps'):
        if isinstance(radiance, (str, EmissionLut)):
            super().__init__(radiance)
            return
        from scipy.interpolate import interp1d
        lut_random = np.linspace(0.0, 1.0, n)
        radiance = np.asarray(radiance, dtype=np.float64)
        if costheta is None:
            costheta = np.linspace(0.0, 1.0, radiance.size)
        else:
            costheta = np.asarray(costheta, dtype=np.float64)
        if radiance.size != costheta.size:
            raise ValueError('The sizes of the radiance and costheta array must be equal!')
        radiance_f = interp1d(costheta, radiance*np.sqrt(1.0 - costheta**2))
        ct_range = costheta.min(), costheta.max()
        if meth == 'simps':
            npts = max(int((max(2*n, npts)//2))*2 + 1, 3)
            ct = np.linspace(ct_range[0], ct_range[1], max(n, npts))
            cdf = np.zeros([int(ct.size//2) + 1])
            radiance_ct = radiance_f(ct)
            dx = (ct[-1] - ct[0])/(ct.size - 1)
            cdf[1:] = dx/3.0*(
                radiance_ct[:-2:2] + 4.0*radiance_ct[1:-1:2] + radiance_ct[2::2])
            cdf = cdf.cumsum()
            cdf /= cdf[-1]
            lut = interp1d(cdf, ct[::2])(lut_random)
        elif meth == 'quad':
            from scipy.integrate import quad
            ct = np.linspace(ct_range[0], ct_range[1], max(n, npts))
            cdf = np.zeros_like(ct)
            for index, ct_item in enumerate(ct):
                cdf[index] = quad(radiance_f, ct_range[0], ct_item)[0]
            cdf /= cdf[-1]
            lut = interp1d(cdf, ct)(lut_random)
        else:
            raise ValueError('Unknown integration method "{}"!'.format(meth))
        super().__init__(lut, 0.0, 1.0)
