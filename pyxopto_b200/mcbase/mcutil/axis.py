"""Accumulator axes (mirror of ``xopto/mcbase/mcutil/axis.py:28-260``)."""
import numpy as np


class Axis:
    def __init__(self, start=0.0, stop: float = 1.0, n: int = 1, logscale: bool = False):
        if isinstance(start, Axis):
            other = start
            start, stop, n, logscale = other.start, other.stop, other.n, other.logscale
        start, stop, n, logscale = float(start), float(stop), int(n), bool(logscale)
        self._logscale, self._n = logscale, n
        if logscale:
            start = max(start, float(np.finfo(np.float64).eps))
            self._edges = np.logspace(np.log(start), np.log(stop), n + 1, base=np.e)
            self._step = (np.log(stop) - np.log(start))/n
        else:
            self._edges = np.linspace(start, stop, n + 1)
            self._step = (stop - start)/n
        self._span = np.array((start, stop), dtype=np.float64)
        self._scaled_span = np.log(self._span) if logscale else self._span
        self._centers = 0.5*(self._edges[:-1] + self._edges[1:])

    logscale = property(lambda self: self._logscale)
    step = property(lambda self: self._step)
    n = property(lambda self: self._n)
    span = property(lambda self: self._span)
    start = property(lambda self: self._span[0])
    stop = property(lambda self: self._span[1])
    scaled_span = property(lambda self: self._scaled_span)
    scaled_start = property(lambda self: self._scaled_span[0])
    scaled_stop = property(lambda self: self._scaled_span[1])
    edges = property(lambda self: self._edges)
    centers = property(lambda self: self._centers)

    def todict(self) -> dict:
        return {'start': self.start, 'stop': self.stop, 'n': self.n,
                'logscale': self.logscale, 'type': type(self).__name__}

    @classmethod
    def fromdict(cls, data: dict):
        data = dict(data)
        data.pop('type', None)
        return cls(**data)

    def __repr__(self):
        return '{}(start={}, stop={}, n={}, logscale={})'.format(
            type(self).__name__, self.start, self.stop, self.n, self.logscale)


class RadialAxis(Axis):
    """Radial axis whose bin centers are area-weighted (axis.py:202-249)."""

    def __init__(self, start=0.0, stop: float = 1.0, n: int = 1, logscale: bool = False):
        super().__init__(start, stop, n, logscale)
        e = self._edges
        self._centers = (2.0/3.0)*(e[:-1]**2 + e[:-1]*e[1:] + e[1:]**2)/(e[:-1] + e[1:])


class SymmetricAxis:
    """Bins placed symmetrically around a center, linear or logarithmic
    (axis.py:267-440)."""

    def __init__(self, center=0.0, range: float = 1.0, n_half: int = 1000,
                 logscale: bool = False):
        if isinstance(center, SymmetricAxis):
            a = center
            range, center, n_half, logscale = a.range, a.center, a.n_half, a.logscale
        range, center, n_half, logscale = float(range), float(center), int(n_half), bool(logscale)
        self._range, self._center, self._logscale = range, center, logscale
        self._n_half, self._n = n_half, 2*n_half
        if logscale:
            eps = np.finfo(np.float64).eps
            tmp = center + np.logspace(np.log(eps), np.log(range), n_half + 1, base=np.e)
            right = tmp[1:]
            self._offset = eps
            self._edges = np.hstack((-right[::-1], [center], right))
            self._step = (np.log(range) - np.log(eps))/n_half
        else:
            self._offset = 0.0
            self._edges = np.linspace(center - range, center + range, 2*n_half + 1)
            self._step = (self._edges[-1] - self._edges[0])/self._n
        self._span = np.array((center + self._offset, center + range), dtype=np.float64)
        self._scaled_offset = np.log(self._offset) if logscale else self._offset
        self._centers = 0.5*(self._edges[:-1] + self._edges[1:])

    logscale = property(lambda self: self._logscale)
    center = property(lambda self: self._center)
    range = property(lambda self: self._range)
    step = property(lambda self: self._step)
    n_half = property(lambda self: self._n_half)
    n = property(lambda self: self._n)
    span = property(lambda self: self._span)
    start = property(lambda self: self._span[0])
    stop = property(lambda self: self._span[1])
    scaled_offset = property(lambda self: self._scaled_offset)
    edges = property(lambda self: self._edges)
    centers = property(lambda self: self._centers)

    def todict(self) -> dict:
        return {'range': self._range, 'center': self._center, 'n_half': self._n_half,
                'logscale': self._logscale, 'type': 'SymmetricAxis'}

    @classmethod
    def fromdict(cls, data: dict):
        data = dict(data)
        data.pop('type', None)
        return cls(**data)

    def __repr__(self):
        return 'SymmetricAxis(range={}, center={}, n_half={}, logscale={})'.format(
            self._range, self._center, self._n_half, self._logscale)
