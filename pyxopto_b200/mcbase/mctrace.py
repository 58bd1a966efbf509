"""Photon packet trace + host-side filter (mirror of ``xopto/mcbase/mctrace.py``)."""
import numpy as np

from ..cl import cltypes
from .mcobject import McObject


def _range(v):
    return (float(v[0]), float(v[1]))


class Filter:
    """Selects packets by their *terminal* event (mctrace.py:53-441).
    Each of x, y, z, pz, pl is a ``(low, high)`` tuple or a list of such tuples
    (OR-ed); ``r`` is ``(rmin, rmax, (x0, y0))``; ``dir`` is
    ``(cosmin, cosmax, (px, py, pz))``.  Different keys are AND-ed."""

    def __init__(self, x=None, y=None, z=None, pz=None, r=None, dir=None, pl=None):
        if isinstance(x, Filter):
            f = x
            x, y, z, pz, r, dir, pl = f.x, f.y, f.z, f.pz, f.r, f.dir, f.pl
        self._x = self._ranges('x', x)
        self._y = self._ranges('y', y)
        self._z = self._ranges('z', z)
        self._pz = self._ranges('pz', pz)
        self._pl = self._ranges('pl', pl)
        self._r = self._check_r(r)
        self._dir = self._check_dir(dir)

    @staticmethod
    def _ranges(name, value):
        if value is None:
            return None
        if not isinstance(value, (list, tuple)):
            raise TypeError('Filter parameter {} must be a tuple of two float '
                            'values (low, high), or a list of such tuples!'.format(name))
        if isinstance(value[0], (list, tuple)):
            return [_range(item) for item in value]
        return (_range(value),)

    @staticmethod
    def _check_r(r):
        if r is None:
            return None
        if not isinstance(r, (list, tuple)):
            raise TypeError('Filter parameter r must be a tuple '
                            '(rmin, rmax, (x_origin, y_origin))!')

        def one(item):
            origin = item[2] if len(item) > 2 else (0.0, 0.0)
            return (float(item[0]), float(item[1]), (float(origin[0]), float(origin[1])))
        if isinstance(r[0], (list, tuple)):
            return [one(item) for item in r]
        return (one(r),)

    @staticmethod
    def _check_dir(d):
        if d is None:
            return None

        def one(item):
            p = np.asarray(item[2], dtype=np.float64)
            length = np.linalg.norm(p)
            if length == 0.0:
                raise ValueError('Direction vector length must not be 0!')
            return (float(item[0]), float(item[1]), tuple(p/length))
        if isinstance(d[0], (list, tuple)):
            return [one(item) for item in d]
        return (one(d),)

    x = property(lambda self: self._x)
    y = property(lambda self: self._y)
    z = property(lambda self: self._z)
    pz = property(lambda self: self._pz)
    r = property(lambda self: self._r)
    dir = property(lambda self: self._dir)
    pl = property(lambda self: self._pl)

    def todict(self):
        return {'type': 'Filter', 'x': self._x, 'y': self._y, 'z': self._z,
                'pz': self._pz, 'dir': self._dir, 'r': self._r, 'pl': self._pl}

    @classmethod
    def fromdict(cls, data):
        d = dict(data)
        d.pop('type')
        return cls(**d)

    def mask(self, trace_obj) -> np.ndarray:
        """Boolean mask of packets whose terminal event passes the filter."""
        term = trace_obj.terminal
        valid = np.ones(term.shape[0], dtype=bool)
        var = np.zeros_like(valid)

        def stage(items, fn, reset=True):
            nonlocal valid, var
            if reset:
                var = np.zeros_like(valid)
            for item in items:
                var |= fn(item) & valid
            valid &= var

        if self._x is not None:
            stage(self._x, lambda g: (term['x'] >= g[0]) & (term['x'] <= g[1]))
        if self._y is not None:
            stage(self._y, lambda g: (term['y'] >= g[0]) & (term['y'] <= g[1]))
        if self._z is not None:
            stage(self._z, lambda g: (term['z'] >= g[0]) & (term['z'] <= g[1]))
        if self._pz is not None:
            stage(self._pz, lambda g: (term['pz'] >= g[0]) & (term['pz'] <= g[1]))
        if self._r is not None:
            def in_r(g):
                dx, dy = term['x'] - g[2][0], term['y'] - g[2][1]
                rr = dx**2 + dy**2
                return (rr >= g[0]**2) & (rr <= g[1]**2)
            stage(self._r, in_r)
        if self._dir is not None:
            def in_dir(g):
                ct = term['px']*g[2][0] + term['py']*g[2][1] + term['pz']*g[2][2]
                return (ct >= g[0]) & (ct <= g[1])
            stage(self._dir, in_dir)
        if self._pl is not None and trace_obj.plon:
            # the reference does not clear the OR-mask before the pl stage
            # (mctrace.py:300-308); kept for drop-in parity
            stage(self._pl, lambda g: (term['pl'] >= g[0]) & (term['pl'] <= g[1]),
                  reset=False)
        return valid

    def cu_pack(self, plon: bool = True):
        """(counts, ranges) for the device-side filter (TraceFilterFlags in
        csrc/kernels/xo_trace_kernels.cuh): ``counts`` = uint32[8] {nx, ny, nz,
        npz, nr, ndir, npl, plon}, ``ranges`` = the float32 constants in that
        order.  Constants are rounded to binary32 exactly where numpy rounds
        them when ``mask`` compares float32 trace fields with Python floats."""
        f32 = np.float32
        vals = []
        counts = []
        for items in (self._x, self._y, self._z, self._pz):
            counts.append(0 if items is None else len(items))
            for g in (items or ()):
                vals += [f32(g[0]), f32(g[1])]
        counts.append(0 if self._r is None else len(self._r))
        for g in (self._r or ()):
            vals += [f32(g[0]**2), f32(g[1]**2), f32(g[2][0]), f32(g[2][1])]
        counts.append(0 if self._dir is None else len(self._dir))
        for g in (self._dir or ()):
            vals += [f32(g[0]), f32(g[1]), f32(g[2][0]), f32(g[2][1]), f32(g[2][2])]
        counts.append(0 if self._pl is None else len(self._pl))
        for g in (self._pl or ()):
            vals += [f32(g[0]), f32(g[1])]
        counts.append(int(bool(plon)))
        return (np.array(counts, dtype=np.uint32),
                np.array(vals + [0.0], dtype=np.float32))

    def __call__(self, trace_obj, update: bool = True):
        if not isinstance(trace_obj, Trace):
            raise TypeError('Trace filter can be applied only to Trace objects!')
        too_long = trace_obj.overflow
        valid = self.mask(trace_obj)
        selected = valid & ~too_long
        n_dropped = int(np.count_nonzero(valid & too_long))
        data, n = trace_obj.data[selected, :], trace_obj.n[selected]
        out = trace_obj if update else Trace(trace_obj)
        out.data, out.n = data, n
        out._terminal = out._overflow_mask = None
        return out, n_dropped

    def __repr__(self):
        return 'Filter(x={}, y={}, z={}, pz={}, dir={}, r={}, pl={})'.format(
            self.x, self.y, self.z, self.pz, self.dir, self.r, self.pl)


class Trace(McObject):
    TRACE_NONE, TRACE_START, TRACE_END, TRACE_ALL = 0, 1, 2, 7
    TRACE_ENTRY_LEN = 8
    TRACE_EVENT_REFLECTION = 1
    TRACE_EVENT_REFRACTION = 2
    TRACE_EVENT_BOUNDARY_HIT = 4
    TRACE_EVENT_LAUNCH = 8
    TRACE_EVENT_ABSORPTION = 16
    TRACE_EVENT_SCATTERING = 32
    TRACE_EVENT_TERMINATION = 64
    TRACE_EVENT_ESCAPE = 128
    TRACE_EVENT_ALL = -1

    cu_type = 'xo::TraceCfg'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClTrace(cltypes.Structure):
            _fields_ = [('max_events', T.mc_int_t), ('data_buffer_offset', T.mc_size_t),
                        ('count_buffer_offset', T.mc_size_t), ('event_mask', T.mc_uint_t)]
        return ClTrace

    def cl_options(self, mc):
        return [('MC_USE_EVENTS', self._event_mask is not None),
                ('MC_USE_TRACE', self._options),
                ('TRACE_ENTRY_LEN', int(Trace.TRACE_ENTRY_LEN)),
                ('MC_USE_SAMPLING_VOLUME', True),
                ('MC_TRACK_OPTICAL_PATHLENGTH', bool(self.plon))]

    def __init__(self, maxlen=500, options: int = TRACE_ALL, filter: Filter = None,
                 event_mask: int = None, plon: bool = True):
        super().__init__()
        if isinstance(maxlen, Trace):
            t = maxlen
            self._maxlen, self._data, self._n = t.maxlen, t.data, t.n
            self._options, self._filter = t.options, t.filter
            self._n_dropped, self._plon, self._event_mask = t.dropped, t.plon, t.event_mask
        else:
            self._data = self._n = None
            self._options = int(options) & 7
            self.maxlen = maxlen
            self._filter = filter
            self._n_dropped = 0
            self._plon = bool(plon)
            self._event_mask = None if event_mask is None else int(event_mask)
        self._terminal = self._overflow_mask = None

    def _set_maxlen(self, maxlen):
        if self._options == Trace.TRACE_ALL:
            pass
        elif self._options == (Trace.TRACE_START | Trace.TRACE_END):
            maxlen = 2
        else:
            maxlen = 1
        if maxlen < 1:
            raise ValueError('Maximum trace length must be at least 1!')
        self._maxlen = int(maxlen)

    maxlen = property(lambda self: self._maxlen, _set_maxlen)
    options = property(lambda self: self._options)
    plon = property(lambda self: self._plon)
    event_mask = property(lambda self: self._event_mask)
    dropped = property(lambda self: self._n_dropped)

    def _set_filter(self, f):
        if not isinstance(f, Filter):
            raise TypeError('Expected a "Filter" instance but got "{}"'.format(type(f)))
        self._filter = f

    filter = property(lambda self: self._filter, _set_filter)

    # The rows of a device-filtered run may still be on the device only: `_data`
    # then fetches them on first access (Mc.run installs `_lazy_rows`; the
    # simulator materialises an outstanding result before it reuses the buffer).
    # A result that is only handed to Mc.sampling_volume never downloads its rows.
    def _get_rows(self):
        loader = self.__dict__.get('_lazy_rows')
        if loader is not None:
            self.__dict__['_lazy_rows'] = None
            self.__dict__['_rows'] = loader()
        return self.__dict__.get('_rows')

    def _set_rows(self, rows):
        self.__dict__['_lazy_rows'] = None
        self.__dict__['_rows'] = rows

    _data = property(_get_rows, _set_rows)

    # token of the device-resident copy of the rows (set by Mc.run, dropped as
    # soon as the host arrays are replaced): lets Mc.sampling_volume skip the
    # re-upload of rows that never left the device
    _device_token = None

    def _set_data(self, d):
        self._data = d
        self._device_token = None

    def _set_lazy_rows(self, loader):
        """Rows that are fetched from the device on first access of ``data``."""
        self.__dict__['_rows'] = None
        self.__dict__['_lazy_rows'] = loader

    rows_on_device_only = property(
        lambda self: self.__dict__.get('_lazy_rows') is not None, None, None,
        'True while the rows of this result have not been downloaded yet.')

    def _set_n(self, n):
        self._n = n
        self._device_token = None

    data = property(lambda self: self._data, _set_data)
    n = property(lambda self: self._n, _set_n)
    nphotons = property(lambda self: 0 if self._n is None else self._n.size)

    @property
    def terminal(self):
        if self._terminal is None and self._data is not None:
            last = np.minimum(self.n - 1, self.maxlen - 1)
            self._terminal = self._data[np.arange(self.n.size), last]
        return self._terminal

    @property
    def overflow(self):
        if self._overflow_mask is None and self._data is not None:
            self._overflow_mask = self.n >= self.maxlen
        return self._overflow_mask

    def dtype(self, mc=None) -> np.dtype:
        f = np.float32 if mc is None else mc.types.np_float
        return np.dtype([(k, f) for k in ('x', 'y', 'z', 'px', 'py', 'pz', 'w', 'pl')])

    def float_len(self, nphotons: int) -> int:
        return Trace.TRACE_ENTRY_LEN*self.maxlen*nphotons

    def np_buffer(self, mc, allocation, nphotons=None, **kwargs):
        if allocation.dtype == np.dtype(mc.types.np_int):
            return np.empty((nphotons,), dtype=mc.types.np_int)
        if allocation.dtype == np.dtype(mc.types.np_float):
            return np.empty((nphotons, self._maxlen), dtype=self.dtype(mc))
        raise RuntimeError('Unexpected buffer allocation!')

    def cl_pack(self, mc, target=None, nphotons: int = None):
        if target is None:
            target = self.cl_type(mc)()
        if nphotons is None:
            raise ValueError('The number of photon packets was not defined!')
        target.data_buffer_offset = mc.cl_allocate_rw_float_buffer(
            self, (Trace.TRACE_ENTRY_LEN*self.maxlen*nphotons,)).offset
        target.count_buffer_offset = mc.cl_allocate_rw_int_buffer(
            self, (nphotons,)).offset
        target.max_events = self.maxlen
        target.event_mask = 0 if self._event_mask is None else self._event_mask & 0xFFFFFFFF
        return target

    def apply_filter(self):
        if self._filter is not None:
            self._n_dropped = self._filter(self, update=True)[1]
            self._terminal = None

    def update_data(self, mc, data, nphotons, prefiltered: bool = False,
                    n_dropped: int = 0, **kwargs):
        """Adopt the rows of a run (mctrace.py:1120-1172).  ``prefiltered``: the
        rows were already selected by the device-side filter, which also counted
        ``n_dropped``; the host filter is then skipped."""
        new_data = data[np.dtype(mc.types.np_float)][0]
        new_n = data[np.dtype(mc.types.np_int)][0]
        self._terminal = self._overflow_mask = None
        if self._data is not None:
            if self._data.shape[1] != new_data.shape[1]:
                raise ValueError('Cannot update the trace with trace data '
                                 'of different maximum length!')
            tmp = type(self)(self)
            tmp.data = tmp.n = None
            tmp.update_data(mc, data, nphotons, prefiltered=prefiltered,
                            n_dropped=n_dropped)
            self._n = np.hstack([self._n, tmp.n])
            self._data = np.vstack([self._data, tmp.data])
            self._device_token = None
        else:
            self._n, self._data = new_n, new_data
            if prefiltered:
                self._n_dropped = int(n_dropped)
            else:
                self.apply_filter()

    def todict(self):
        return {'type': 'Trace', 'maxlen': self._maxlen, 'options': self._options,
                'plon': self._plon,
                'filter': None if self._filter is None else self._filter.todict()}

    def __len__(self):
        return self.nphotons
