"""Kernel data type families (mirror of ``xopto/mcbase/mctypes.py:542-1044``).

The photon-packet kernels of this engine compute in fp32 with 32-bit integers /
size_t and 64-bit fixed-point accumulators - the reference's default family
``McDataTypesSingle``.  ``McDataTypesSingleCnt64`` only widens the host-side
packet counter bookkeeping (runs above 2**32-1 packets are split in batches).
``McDataTypesDouble`` / ``McDataTypesDoubleCnt64`` (mctypes.py:647-748,991-1044)
switch ``mc_fp_t`` to binary64: packed structs, trace rows and lookup tables carry
doubles and the kernels are compiled with ``XO_DOUBLE`` (reference expression order,
IEEE double arithmetic, CUDA's double-precision elementary functions).
"""
import ctypes

import numpy as np

from ..cl import cltypes


class _Vec(cltypes.Structure):
    def fromarray(self, array):
        flat = np.asarray(array).ravel()
        for (name, ctype), value in zip(self._fields_, flat):
            integer = ctype not in (ctypes.c_float, ctypes.c_double)
            setattr(self, name, int(value) if integer else float(value))
        return self

    def toarray(self) -> np.ndarray:
        return np.array([getattr(self, name) for name, _ in self._fields_])

    def tolist(self):
        return self.toarray().tolist()


def _vec(name, ctype, fields):
    return type(name, (_Vec,), {'_fields_': [(f, ctype) for f in fields]})


class _Matrix3(cltypes.Structure):
    def fromarray(self, array):
        flat = np.asarray(array, dtype=np.float64).reshape(9)
        for (name, _), value in zip(self._fields_, flat):
            setattr(self, name, value)
        return self

    def toarray(self) -> np.ndarray:
        return np.array([getattr(self, n) for n, _ in self._fields_]).reshape(3, 3)


_M3_FIELDS = ['a_11', 'a_12', 'a_13', 'a_21', 'a_22', 'a_23', 'a_31', 'a_32', 'a_33']


class McDataTypesSingle:
    """fp32 / int32 / 32-bit size_t / 32-bit packet counter / 64-bit accumulators."""
    mc_fp_t = ctypes.c_float
    mc_int_t = ctypes.c_int32
    mc_uint_t = ctypes.c_uint32
    mc_size_t = ctypes.c_uint32
    mc_cnt_t = ctypes.c_uint32
    mc_accu_t = ctypes.c_uint64

    np_float = np.float32
    np_int = np.int32
    np_uint = np.uint32
    np_size = np.uint32
    np_cnt = np.uint32
    np_accu = np.uint64

    mc_fp_maxint = 0x7FFFFF
    mc_accu_k = 0x7FFFFF                # MC_INT_ACCUMULATOR_K (mctypes.py:500-503)
    mc_cnt_max = 0xFFFFFFFF
    mc_accu_max = 0xFFFFFFFFFFFFFFFF
    eps = float(np.finfo(np.float32).eps)

    mc_point2f_t = _vec('mc_point2f_t', ctypes.c_float, 'xy')
    mc_point3f_t = _vec('mc_point3f_t', ctypes.c_float, 'xyz')
    mc_point4f_t = _vec('mc_point4f_t', ctypes.c_float, 'xyzw')
    mc_point2_t = _vec('mc_point2_t', ctypes.c_int32, 'xy')
    mc_point3_t = _vec('mc_point3_t', ctypes.c_int32, 'xyz')
    mc_point4_t = _vec('mc_point4_t', ctypes.c_int32, 'xyzw')
    mc_point2s_t = _vec('mc_point2s_t', ctypes.c_uint32, 'xy')
    mc_point3s_t = _vec('mc_point3s_t', ctypes.c_uint32, 'xyz')
    mc_point4s_t = _vec('mc_point4s_t', ctypes.c_uint32, 'xyzw')
    mc_matrix3f_t = type('mc_matrix3f_t', (_Matrix3,),
                         {'_fields_': [(f, ctypes.c_float) for f in _M3_FIELDS]})
    mc_matrix2f_t = _vec('mc_matrix2f_t', ctypes.c_float, ['a_11', 'a_12', 'a_21', 'a_22'])

    @classmethod
    def cl_options(cls, *_):
        return [('MC_USE_DOUBLE_PRECISION', False),
                ('MC_USE_64_BIT_SIZE_T', False), ('MC_USE_64_BIT_INTEGER', False),
                ('MC_USE_64_BIT_PACKET_COUNTER', False),
                ('MC_USE_64_BIT_ACCUMULATORS', True),
                ('MC_INT_ACCUMULATOR_K', cls.mc_accu_k)]


class McDataTypesSingleCnt64(McDataTypesSingle):
    """As above; the host splits runs of more than 2**32-1 packets in batches."""
    np_cnt = np.uint64
    mc_cnt_max = 0xFFFFFFFFFFFFFFFF


class McDataTypesDouble(McDataTypesSingle):
    """fp64 / int32 / 32-bit size_t / 32-bit packet counter / 64-bit accumulators."""
    mc_fp_t = ctypes.c_double
    np_float = np.float64
    mc_fp_maxint = 0xFFFFFFFFFFFFF          # 52 bits (McDouble.mc_fp_maxint)
    eps = float(np.finfo(np.float64).eps)

    mc_point2f_t = _vec('mc_point2f_t', ctypes.c_double, 'xy')
    mc_point3f_t = _vec('mc_point3f_t', ctypes.c_double, 'xyz')
    mc_point4f_t = _vec('mc_point4f_t', ctypes.c_double, 'xyzw')
    mc_matrix3f_t = type('mc_matrix3f_t', (_Matrix3,),
                         {'_fields_': [(f, ctypes.c_double) for f in _M3_FIELDS]})
    mc_matrix2f_t = _vec('mc_matrix2f_t', ctypes.c_double, ['a_11', 'a_12', 'a_21', 'a_22'])

    @classmethod
    def cl_options(cls, *_):
        return [('MC_USE_DOUBLE_PRECISION', True),
                ('MC_USE_64_BIT_SIZE_T', False), ('MC_USE_64_BIT_INTEGER', False),
                ('MC_USE_64_BIT_PACKET_COUNTER', False),
                ('MC_USE_64_BIT_ACCUMULATORS', True),
                ('MC_INT_ACCUMULATOR_K', cls.mc_accu_k)]


class McDataTypesDoubleCnt64(McDataTypesDouble):
    """As above; the host splits runs of more than 2**32-1 packets in batches."""
    np_cnt = np.uint64
    mc_cnt_max = 0xFFFFFFFFFFFFFFFF


McDataTypes = McDataTypesSingle
