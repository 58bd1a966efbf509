"""Optical fiber description (mirror of ``xopto/mcbase/mcutil/fiber.py`` MultimodeFiber)."""


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
