"""Builds pyxopto_b200/data/safeprimes_a_500k.npz with tools/gen_safeprimes.c.

Optionally cross-checks against the reference's shipped table when
/root/reference is present (this container only).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'pyxopto_b200', 'data', 'safeprimes_a_500k.npz')
N = 500000


def main():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'gen_safeprimes')
        subprocess.check_call(['gcc', '-O2', os.path.join(ROOT, 'tools', 'gen_safeprimes.c'),
                               '-o', exe])
        raw = os.path.join(tmp, 'a.bin')
        subprocess.check_call([exe, str(N), raw])
        a = np.fromfile(raw, dtype=np.uint32)
    assert a.size == N
    ref = '/root/reference/xopto/data/primes/safeprimes_base32_500k.npz'
    if os.path.exists(ref):
        ref_a = np.load(ref)['data'][:, 0].astype(np.uint32)
        print('matches the reference table:', bool(np.array_equal(a, ref_a)))
        assert np.array_equal(a, ref_a)
    # store first value + negative deltas (uint16 fits: gaps are < 65536) for size
    gaps = (a[:-1].astype(np.int64) - a[1:].astype(np.int64))
    assert gaps.min() > 0
    if gaps.max() < 65536:
        np.savez_compressed(OUT, first=np.uint32(a[0]), gaps=gaps.astype(np.uint16))
    else:
        np.savez_compressed(OUT, first=np.uint32(a[0]), gaps=gaps.astype(np.uint32))
    print('wrote', OUT, os.path.getsize(OUT), 'bytes; max gap', gaps.max())


if __name__ == '__main__':
    sys.exit(main())
