"""Voxel grid (mirror of ``xopto/mcvox/mcgeometry/voxel.py``).

``Voxels.material`` is the int32 material-index array indexed [z, y, x]; the
packed ``McVoxelConfig`` is {top_left, bottom_right, size (3 fp each), shape (3 int)}.
"""
from typing import Tuple

import numpy as np

from ..cl import cltypes
from ..mcbase.mcobject import McObject
from ..mcbase.mcutil.axis import Axis  # noqa: F401


class Voxel(McObject):
    @staticmethod
    def cl_type(mc):
        class ClVoxel(cltypes.Structure):
            _fields_ = [('material_index', mc.types.mc_int_t)]
        return ClVoxel

    @staticmethod
    def np_dtype(mc=None):
        return np.dtype([('material_index', np.int32)])


class Voxels(McObject):
    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClVoxels(cltypes.Structure):
            _fields_ = [('top_left', T.mc_point3f_t), ('bottom_right', T.mc_point3f_t),
                        ('size', T.mc_point3f_t), ('shape', T.mc_point3_t)]
        return ClVoxels

    def __init__(self, xaxis: Axis, yaxis: Axis, zaxis: Axis, voxel=None):
        super().__init__()
        if xaxis.logscale or yaxis.logscale or zaxis.logscale:
            raise ValueError('Logarithmic scale is not supported!')
        self._x_axis, self._y_axis, self._z_axis = xaxis, yaxis, zaxis
        self._voxel_size = (abs(xaxis.step), abs(yaxis.step), abs(zaxis.step))
        self._top_left = (min(xaxis.span), min(yaxis.span), min(zaxis.span))
        self._bottom_right = (max(xaxis.span), max(yaxis.span), max(zaxis.span))
        self._voxel_type = Voxel if voxel is None else voxel
        self._data = None
        self._grid = None
        self._update = True

    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    zaxis = property(lambda self: self._z_axis)
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    dx = property(lambda self: self._voxel_size[0])
    dy = property(lambda self: self._voxel_size[1])
    dz = property(lambda self: self._voxel_size[2])
    voxel_size = property(lambda self: self._voxel_size)
    top_left = property(lambda self: self._top_left)
    shape = property(lambda self: (self._z_axis.n, self._y_axis.n, self._x_axis.n))
    size = property(lambda self: self._z_axis.n*self._y_axis.n*self._x_axis.n)
    grid = property(lambda self: self.meshgrid())

    def data(self, mc=None) -> np.ndarray:
        if self._data is None:
            self._data = np.zeros(self.shape, dtype=self._voxel_type.np_dtype(mc))
        return self._data

    material = property(lambda self: self.data()['material_index'], None, None,
                        '3D array of material indices [z, y, x].')

    def __getitem__(self, item):
        return self.data()[item]

    def __setitem__(self, item, value):
        self.data()[item] = value

    def index(self, position) -> Tuple[int, int, int]:
        return tuple(int((position[i] - self._top_left[i])//self._voxel_size[i])
                     for i in range(3))

    def center(self, index) -> Tuple[float, float, float]:
        return tuple(self._top_left[i] + (index[i] + 0.5)*self._voxel_size[i]
                     for i in range(3))

    def isvalid(self, index) -> bool:
        return (self.shape[2] > index[0] >= 0) and (self.shape[1] > index[1] >= 0) and \
               (self.shape[0] > index[2] >= 0)

    def contains(self, position) -> bool:
        return self._x_axis.start <= position[0] < self._x_axis.stop and \
               self._y_axis.start <= position[1] < self._y_axis.stop and \
               self._z_axis.start <= position[2] < self._z_axis.stop

    def intersect(self, pos, dir):
        """Entry point + surface normal of a ray with the grid box (voxel.py:317-378)."""
        invdir = [np.inf, np.inf, np.inf]
        for ax in (0, 1, 2):
            if dir[ax] != 0.0:
                invdir[ax] = 1.0/dir[ax]
        xmin, xmax = self._x_axis.edges[[0, -1]]
        ymin, ymax = self._y_axis.edges[[0, -1]]
        zmin, zmax = self._z_axis.edges[[0, -1]]
        t1, t2 = (xmin - pos[0])*invdir[0], (xmax - pos[0])*invdir[0]
        t3, t4 = (ymin - pos[1])*invdir[1], (ymax - pos[1])*invdir[1]
        t5, t6 = (zmin - pos[2])*invdir[2], (zmax - pos[2])*invdir[2]
        txmin, txmax = min(t1, t2), max(t1, t2)
        tymin, tymax = min(t3, t4), max(t3, t4)
        tzmin, tzmax = min(t5, t6), max(t5, t6)
        tmin, tmax = max(txmin, tymin, tzmin), min(txmax, tymax, tzmax)
        if tmax < tmin or tmax < 0.0:
            return None, None
        t = tmin if tmin >= 0.0 else tmax
        intersection = (pos[0] + t*dir[0], pos[1] + t*dir[1], pos[2] + t*dir[2])
        normal = [0.0, 0.0, 0.0]
        normal[0] = int(txmin >= tymin and txmin >= tzmin)*np.sign(dir[0])
        normal[1] = int(not normal[0] and tymin >= tzmin)*np.sign(dir[1])
        normal[2] = int(not (normal[0] or normal[1]))*np.sign(dir[2])
        return intersection, normal

    def update_required(self, clear: bool = True) -> bool:
        value = self._update
        if clear:
            self._update = False
        return value

    def update(self):
        self._update = True

    def meshgrid(self):
        if self._grid is None:
            self._grid = np.meshgrid(self._z_axis.centers, self._y_axis.centers,
                                     self._x_axis.centers, indexing='ij')
        return self._grid

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.top_left.fromarray(self._top_left)
        target.bottom_right.fromarray(self._bottom_right)
        target.size.fromarray(self._voxel_size)
        target.shape.x, target.shape.y, target.shape.z = \
            self._x_axis.n, self._y_axis.n, self._z_axis.n
        return target

    def todict(self):
        return {'type': 'Voxels', 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'zaxis': self._z_axis.todict()}
