"""Optical fiber description (mirror of ``xopto/mcbase/mcutil/fiber.py`` MultimodeFiber,
FiberLayout)."""
import numpy as np


class MultimodeFiber:
    def __init__(self, dcore, dcladding: float = None, ncore: float = None,
                 na: float = None):
        if isinstance(dcore, MultimodeFiber):
            f = dcore
            dcore, dcladding, ncore, na = f.dcore, f.dcladding, f.ncore, f.na
        self.dcore = float(dcore)
        self.dcladding = float(dcladding)
        self.ncore = float(ncore)
        self.na = float(na)

    @staticmethod
    def compute_na(ncore: float, ncladding: float) -> float:
        return (ncore**2 - ncladding**2)**0.5

    @staticmethod
    def compute_ncladding(ncore: float, na: float) -> float:
        return (ncore**2 - na**2)**0.5

    ncladding = property(lambda self: self.compute_ncladding(self.ncore, self.na), None, None,
                         'Refractive index of the fiber cladding.')

    def todict(self) -> dict:
        return {'dcore': self.dcore, 'dcladding': self.dcladding,
                'ncore': self.ncore, 'na': self.na, 'type': 'MultimodeFiber'}

    @classmethod
    def fromdict(cls, data: dict):
        data = dict(data)
        data.pop('type', None)
        return cls(**data)

    def __repr__(self):
        return 'MultimodeFiber(dcore={}, dcladding={}, ncore={}, na={})'.format(
            self.dcore, self.dcladding, self.ncore, self.na)


class MultimodeFiberLut:
    """Multimode fiber with tabulated emission / collection characteristics
    (mcutil/fiber.py:224-382)."""
    compute_na = staticmethod(MultimodeFiber.compute_na)
    compute_ncladding = staticmethod(MultimodeFiber.compute_ncladding)

    def __init__(self, dcore: float, dcladding: float, ncore: float, ncladding: float = None,
                 emission=None, collection=None):
        from .lut import CollectionLut, EmissionLut
        self.dcore = float(dcore)
        self.dcladding = float(dcladding)
        self.ncore = float(ncore)
        self._ncladding = self.ncore if ncladding is None else float(ncladding)
        if isinstance(emission, str):
            emission = EmissionLut.fromfile(emission)
        elif isinstance(emission, np.ndarray):
            emission = EmissionLut(emission)
        if isinstance(collection, str):
            collection = CollectionLut.fromfile(collection)
        self._emission_lut = emission
        self._collection_lut = collection

    def validate(self):
        if self.dcore > self.dcladding:
            raise ValueError('Fiber core diameter must be smaller than the fiber '
                             'cladding diameter!')

    ncladding = property(lambda self: self._ncladding)
    emission = property(lambda self: self._emission_lut)
    collection = property(lambda self: self._collection_lut)

    def todict(self) -> dict:
        return {'dcore': self.dcore, 'dcladding': self.dcladding, 'ncore': self.ncore,
                'emission': self._emission_lut, 'collection': self._collection_lut,
                'type': type(self).__name__}

    def __repr__(self):
        return 'MultimodeFiberLut(dcore={:f}, dcladding={:f}, ncore={:f})'.format(
            self.dcore, self.dcladding, self.ncore)


class FiberLayout:
    """A fiber placed at ``position`` (x, y[, z]) and tilted into ``direction``
    (mcutil/fiber.py:384-470)."""
    def __init__(self, fiber, position=(0.0, 0.0, 0.0), direction=(0.0, 0.0, 1.0)):
        if isinstance(fiber, FiberLayout):
            o = fiber
            fiber, position, direction = o.fiber, o.position, o.direction
        self._fiber = fiber
        self._position = np.zeros((3,))
        self._direction = np.array((0.0, 0.0, 1.0))
        self.position = position
        self.direction = direction

    def _set_fiber(self, fiber):
        self._fiber = fiber

    fiber = property(lambda self: self._fiber, _set_fiber)

    def _set_position(self, value):
        value = np.asarray(value, dtype=np.float64).reshape(-1)
        if value.size < 3:
            self._position[:value.size] = value
        else:
            self._position[:] = value[:3]

    position = property(lambda self: self._position, _set_position)

    def _set_direction(self, direction):
        self._direction[:] = direction
        norm = np.linalg.norm(self._direction)
        if norm == 0.0:
            raise ValueError('Direction vector norm/length must not be 0!')
        self._direction *= 1.0/norm

    direction = property(lambda self: self._direction, _set_direction)

    def todict(self) -> dict:
        return {'type': 'FiberLayout', 'fiber': self._fiber.todict(),
                'position': self._position.tolist(), 'direction': self._direction.tolist()}

    @classmethod
    def fromdict(cls, data: dict):
        data = dict(data)
        data.pop('type', None)
        return cls(MultimodeFiber.fromdict(data.pop('fiber')), **data)

    def __repr__(self):
        return 'FiberLayout(fiber={}, position=({}, {}, {}), direction=({}, {}, {}))'.format(
            self._fiber, *self._position, *self._direction)
