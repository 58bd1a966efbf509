"""Sampling volume (mirror of ``xopto/mcbase/mcsv.py``): a voxel grid that
accumulates, for every traced packet, terminal weight x path length travelled
inside each voxel.  Filled by ``Mc.sampling_volume(trace, sv)`` with the CUDA
kernel ``SamplingVolume`` (``csrc/kernels/xo_sv_kernel.cuh``)."""
import numpy as np

from ..cl import cltypes
from . import mctypes
from .mcobject import McObject
from .mcutil.axis import Axis  # noqa: F401  (re-exported like xopto.mcbase.mcsv.Axis)


class SamplingVolume(McObject):
    cu_type = 'xo::SvCfg'

    @staticmethod
    def cl_type(mc):
        T = mc.types
        class ClSamplingVolume(cltypes.Structure):
            _fields_ = [('top_left', T.mc_point3f_t), ('voxel_size', T.mc_point3f_t),
                        ('shape', T.mc_point3s_t), ('multiplier', T.mc_fp_t),
                        ('offset', T.mc_size_t), ('k', T.mc_int_t)]
        return ClSamplingVolume

    def __init__(self, xaxis, yaxis: Axis = None, zaxis: Axis = None):
        data, weight = None, 0.0
        if isinstance(xaxis, SamplingVolume):
            sv = xaxis
            xaxis, yaxis, zaxis = Axis(sv.xaxis), Axis(sv.yaxis), Axis(sv.zaxis)
            if sv.data is not None:         # (reading `data` collects a device-resident grid)
                data = np.copy(sv.data)
            weight = sv.weight
        self._x_axis, self._y_axis, self._z_axis = xaxis, yaxis, zaxis
        self._data, self._weight = data, weight
        self._k = mctypes.McDataTypesSingle.mc_fp_maxint    # McFloat32.mc_fp_maxint
        if self._x_axis.n*self._y_axis.n*self._z_axis.n <= 0:
            raise ValueError('Sampling volume accumulator array has one or '
                             'dimensions equal to zero!')

    shape = property(lambda self: (self._z_axis.n, self._y_axis.n, self._x_axis.n))
    x = property(lambda self: self._x_axis.centers)
    y = property(lambda self: self._y_axis.centers)
    z = property(lambda self: self._z_axis.centers)
    xaxis = property(lambda self: self._x_axis)
    yaxis = property(lambda self: self._y_axis)
    zaxis = property(lambda self: self._z_axis)
    k = property(lambda self: self._k)

    # Device-resident accumulation (``Mc.lazy_sampling_volume``): the simulator keeps
    # the integer grid of this object on the device across ``sampling_volume`` calls
    # and hands over a loader; the grid comes to the host (converted once, from the
    # exact integer sum) when ``data`` is read.
    _pending = None

    def _set_pending(self, loader):
        self._pending = loader

    def _materialize(self):
        loader, self._pending = self._pending, None
        if loader is not None:
            loader(self)

    def _get_data(self):
        self._materialize()
        return self._data

    def _set_data(self, data):
        self._materialize()
        self._data = data

    def _set_weight(self, w):
        self._weight = w

    data = property(_get_data, _set_data, None,
                    'Raw sampling volume accumulator data if any.')
    weight = property(lambda self: self._weight, _set_weight, None,
                      'Total weight of the accumulated photon packets.')

    def clear(self):
        self._materialize()
        if self._data is not None:
            self._data.fill(0)

    def _multiplier(self, mc=None) -> float:
        """Scale of weight x path length before the fixed-point conversion:
        the inverse of the smallest voxel edge (mcsv.py:266-273)."""
        l = np.min([self._x_axis.step, self._y_axis.step, self._z_axis.step])
        if l <= 0.0:
            l = 1.0/self._k
        return 1.0/l

    def update_data(self, mc, accumulators, total_weight, **kwargs):
        multiplier = self._multiplier(mc)
        new_data = np.reshape(accumulators[0], self.shape)
        if self._data is not None:
            self._data += new_data*(1.0/(self.k*multiplier))
            self._weight += float(total_weight)/self._k
        else:
            self._data = new_data*(1.0/(self.k*multiplier))
            self._weight = float(total_weight)/self._k

    def update_scaled(self, scaled, total_weight):
        """``update_data`` with ``accumulators*(1/(k*multiplier))`` already computed
        (bit-identically) on the device; ``scaled`` may be reused by the caller."""
        if self._data is not None:
            self._data += np.reshape(scaled, self.shape)
        else:
            self._data = np.array(scaled, dtype=np.float64).reshape(self.shape)
        self.add_weight(total_weight)

    def add_weight(self, total_weight):
        self._weight = (self._weight or 0.0) + float(total_weight)/self._k

    def cl_pack(self, mc, target=None):
        if target is None:
            target = self.cl_type(mc)()
        target.offset = mc.cl_allocate_rw_accumulator_buffer(self, self.shape).offset
        target.voxel_size.x = self._x_axis.step
        target.voxel_size.y = self._y_axis.step
        target.voxel_size.z = self._z_axis.step
        target.top_left.x = self._x_axis.start
        target.top_left.y = self._y_axis.start
        target.top_left.z = self._z_axis.start
        target.shape.x, target.shape.y, target.shape.z = \
            self._x_axis.n, self._y_axis.n, self._z_axis.n
        target.multiplier = self._multiplier(mc)
        target.k = self._k
        return target

    def todict(self):
        return {'type': 'SamplingVolume', 'xaxis': self._x_axis.todict(),
                'yaxis': self._y_axis.todict(), 'zaxis': self._z_axis.todict()}

    @classmethod
    def fromdict(cls, data):
        d = dict(data)
        if d.pop('type') != cls.__name__:
            raise TypeError('Expected data for type "{}"!'.format(cls.__name__))
        return cls(Axis.fromdict(d.pop('xaxis')), Axis.fromdict(d.pop('yaxis')),
                   Axis.fromdict(d.pop('zaxis')))
