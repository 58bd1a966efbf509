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


def rotation_matrix_2d(a, b) -> np.ndarray:
    """2 x 2 rotation of unit vector a onto b as the reference computes it
    (mcutil/geometry.py:33-60): the sine is taken as +sqrt(1 - cos^2), i.e. the
    rotation is always counter-clockwise by the angle between the vectors."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a = a/np.linalg.norm(a)
    b = b/np.linalg.norm(b)
    cos_theta = np.dot(a, b)
    sin_theta = np.sqrt(1.0 - cos_theta**2)
    return np.array([[cos_theta, -sin_theta], [sin_theta, cos_theta]])
