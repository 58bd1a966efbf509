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


def _resample(x, y, xq):
    """Piecewise-linear resampling ``y(x) -> y(xq)`` (``x`` ascending or descending,
    ``xq`` inside its range).  The segment is chosen with a left-sided search and
    evaluated in the slope form ``y0 + (xq - x0)*(y1 - y0)/(x1 - x0)``, which is the
    arithmetic scipy's ``interp1d`` performs - the reference builds its tables with
    it (mcutil/lut.py:285-362), and the packed float pool has to match to the bit."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    if x.size > 1 and x[0] > x[-1]:
        x, y = x[::-1], y[::-1]
    if xq.size and (xq.min() < x[0] or xq.max() > x[-1]):
        raise ValueError('A value in x_new is outside of the interpolation range.')
    hi = np.clip(np.searchsorted(x, xq), 1, x.size - 1)
    lo = hi - 1
    return (y[hi] - y[lo])/(x[hi] - x[lo])*(xq - x[lo]) + y[lo]


def _cosine_axis(costheta, size: int) -> np.ndarray:
    if costheta is None:
        return np.linspace(0.0, 1.0, size)
    return np.asarray(costheta, dtype=np.float64)


class CollectionLut(LinearLut):
    """Angular sensitivity of a detector over the cosine of the incidence angle,
    tabulated at ``n`` equidistant cosines (counterpart of mcutil/lut.py:327-362)."""
    def __init__(self, sensitivity, costheta=None, n: int = 1000):
        if isinstance(sensitivity, (str, LinearLut)):
            super().__init__(sensitivity)
            return
        values = np.asarray(sensitivity, dtype=np.float64)
        axis = _cosine_axis(costheta, values.size)
        grid = np.linspace(axis.min(), axis.max(), n)
        super().__init__(_resample(axis, values, grid), grid[0], grid[-1])


class EmissionLut(LinearLut):
    """Inverse-CDF table of a source's emission-angle cosine for an azimuthally
    symmetric angular radiance: entry ``u*(n - 1)`` of the table is the cosine at
    which the cumulative emitted power reaches the fraction ``u`` (counterpart of
    mcutil/lut.py:240-325).  The power density over the cosine is
    ``radiance*sin(theta)``; its running integral is taken panel by panel with
    Simpson's rule on an odd number of equidistant nodes (``meth='simps'``) or
    with adaptive quadrature from the lower limit (``meth='quad'``)."""
    def __init__(self, radiance, costheta=None, n: int = 2000, npts: int = 10000,
                 meth: str = 'simps'):
        if isinstance(radiance, (str, LinearLut)):
            super().__init__(radiance)
            return
        if meth not in ('simps', 'quad'):
            raise ValueError('Unknown integration method "{}"!'.format(meth))
        values = np.asarray(radiance, dtype=np.float64)
        axis = _cosine_axis(costheta, values.size)
        if values.size != axis.size:
            raise ValueError('The sizes of the radiance and costheta array must be equal!')
        density = values*np.sqrt(1.0 - axis**2)
        lower, upper = axis.min(), axis.max()
        fractions = np.linspace(0.0, 1.0, n)
        if meth == 'simps':
            nodes = max(n, max(2*(max(2*n, npts)//2) + 1, 3))
            ct = np.linspace(lower, upper, nodes)
            f = _resample(axis, density, ct)
            h = (ct[-1] - ct[0])/(ct.size - 1)
            # one Simpson panel per pair of intervals; the CDF lives on the even nodes
            panels = h/3.0*(f[:-2:2] + 4.0*f[1:-1:2] + f[2::2])
            cdf = np.concatenate(([0.0], panels)).cumsum()
            knots = ct[::2]
        else:
            from scipy.integrate import quad
            knots = np.linspace(lower, upper, max(n, npts))
            cdf = np.array([quad(lambda c: float(_resample(axis, density, c)), lower, c)[0]
                            for c in knots])
        cdf /= cdf[-1]
        super().__init__(_resample(cdf, knots, fractions), 0.0, 1.0)
