"""Helper kernels: RNG known-answer hook (``Mc.rng_test``, cf. mc.py:1320-1362)
and the elementary-function probe used by the math parity tests."""
import numpy as np

_AUX_SRC = '#define XO_DETERMINISTIC {det}\n#include "xo_aux_kernels.cuh"\n'

MATH_FN = {'log': 0, 'sincos': 1, 'cbrt': 2, 'pow': 3, 'exp': 4, 'atan2': 5,
           'sqrt': 6, 'div': 7}


def _aux_module(worker, deterministic: bool):
    return worker._module(_AUX_SRC.format(det=int(deterministic)), deterministic)


def rng_test(worker, n: int, x: int, a: int) -> np.ndarray:
    mod = _aux_module(worker, True)
    out = np.zeros(n, dtype=np.float32)
    buf = worker._buffer('aux_out0', out.nbytes)
    mod.kernel('RngKernel').launch(
        worker._stream, 1, 32, [np.uint64(x), np.uint32(a), np.uint32(n), buf])
    buf.download(worker._stream, out)
    return out


def math_probe(worker, fn: str, in0, in1=None, deterministic: bool = True):
    in0 = np.ascontiguousarray(in0, dtype=np.float32)
    in1 = np.ascontiguousarray(in0 if in1 is None else in1, dtype=np.float32)
    n = in0.size
    mod = _aux_module(worker, deterministic)
    b0 = worker._buffer('aux_in0', in0.nbytes)
    b1 = worker._buffer('aux_in1', in1.nbytes)
    o0 = worker._buffer('aux_out0', in0.nbytes)
    o1 = worker._buffer('aux_out1', in0.nbytes)
    b0.upload(worker._stream, in0)
    b1.upload(worker._stream, in1)
    o0.fill(worker._stream, 0, np.uint32, count=n)
    o1.fill(worker._stream, 0, np.uint32, count=n)
    mod.kernel('MathProbe').launch(
        worker._stream, (n + 255)//256, 256,
        [np.int32(MATH_FN[fn]), np.uint32(n), b0, b1, o0, o1])
    out0, out1 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    o0.download(worker._stream, out0)
    o1.download(worker._stream, out1)
    return out0, out1
