"""Rotation helpers (mirror of ``xopto/mcbase/mcutil/geometry.py:60-133``)."""
import numpy as np


def rotation_matrix(a, b) -> np.ndarray:
    """Matrix R with R @ a/|a| == b/|b| (Rodrigues; undefined for a == -b)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a = a/np.linalg.norm(a)
    b = b/np.linalg.norm(b)
    v = np.cross(a, b)
    c = np.dot(a.flat, b.flat)
    vx = np.array([[0.0, -v[2], v[1]],
                   [v[2], 0.0, -v[0]],
                   [-v[1], v[0], 0.0]], dtype=np.float64)
    return np.identity(3) + vx + np.dot(vx, vx)*(1.0/(1.0 + c))


def transform_base(vfrom, vto) -> np.ndarray:
    return rotation_matrix(vfrom, vto)
