"""Host-side Fresnel/Snell helpers used while packing sources
(mirror of ``xopto/mcbase/mcutil/boundary.py``)."""
import numpy as np


def cos_critical(n1: float, n2: float) -> float:
    return (1.0 - (n2/n1)**2)**0.5 if n1 > n2 else 0.0


def reflectance(n1: float, n2: float, costheta: float = 1.0) -> float:
    if costheta < 0.0:
        raise ValueError('The incidence angle cosine must not be negative!')
    n1, n2 = float(n1), float(n2)
    sintheta = (1.0 - costheta**2)**0.5
    if n1 > n2 and sintheta >= n2/n1:
        return 1.0
    root = (1.0 - (n1/n2*sintheta)**2)**0.5
    a1, a2 = n1*costheta, n2*root
    rs = np.abs((a1 - a2)/(a1 + a2))**2
    b1, b2 = n1*root, n2*costheta
    rp = np.abs((b1 - b2)/(b1 + b2))**2
    return (rs + rp)*0.5


def refract(direction, normal, n1: float, n2: float) -> np.ndarray:
    direction = np.asarray(direction, dtype=np.float64)
    normal = np.asarray(normal, dtype=np.float64)
    dlen, nlen = np.linalg.norm(direction), np.linalg.norm(normal)
    cos1 = np.dot(direction, normal)/(dlen*nlen)
    if abs(cos1) < cos_critical(n1, n2):
        raise ValueError('Cannot refract, the incidence angle '
                         'exceeds the critical angle!')
    n12 = n1/n2
    sin2_squared = n12*n12*(1.0 - cos1*cos1)
    k = n12*abs(cos1) - np.sqrt(1.0 - sin2_squared)
    return n12*direction/dlen - np.sign(cos1)*k*normal/nlen


def reflect(direction, normal) -> np.ndarray:
    direction = np.asarray(direction, dtype=np.float64)
    normal = np.asarray(normal, dtype=np.float64)
    dlen, nlen = np.linalg.norm(direction), np.linalg.norm(normal)
    if dlen > 0.0:
        direction = direction/dlen
    if nlen > 0.0:
        normal = normal/nlen
    return direction - 2.0*np.dot(direction, normal)*normal
