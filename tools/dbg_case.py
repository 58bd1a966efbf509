"""Developer probe: first diverging trace event between the CUDA deterministic
mode and the oracle for one test case with a Trace attached.
    python tools/dbg_case.py mccyl_hg_isopoint_outside [n] [threads]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import cases
import xo_oracle
from helpers import build_sim
from pyxopto_b200.mcbase import mcoptions, mctrace

name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 64
tr = mctrace.Trace(maxlen=400, options=mctrace.Trace.TRACE_ALL, plon=True)
sim, geom, _ = build_sim(name, options=[mcoptions.McDeterministic.on], trace=tr)
sim.run(n, maxthreads=threads, wgsize=64, download=False)
accu, ints, floats = sim.download_raw()
desc = xo_oracle.describe(sim, geom)
ref = xo_oracle.run(desc, n, threads, sim.rng_seeds_x[:threads], sim.rng_seeds_a[:threads],
                    math=xo_oracle.MATH_PORTABLE)
print('accu equal', np.array_equal(accu, ref['accu']), accu[:4], ref['accu'][:4])
P = sim._packed['trace']
ml = 400
f = floats[P.data_buffer_offset:P.data_buffer_offset + n*ml*8].reshape(n, ml, 8)
g = ref['floats'][P.data_buffer_offset:P.data_buffer_offset + n*ml*8].reshape(n, ml, 8)
ci, cg = ints[P.count_buffer_offset:P.count_buffer_offset + n], \
    ref['ints'][P.count_buffer_offset:P.count_buffer_offset + n]
bad = np.nonzero((f.view(np.uint32) != g.view(np.uint32)).any(axis=(1, 2)) | (ci != cg))[0]
print('packets differing', bad.size, 'of', n, 'first', bad[:10])
if bad.size:
    p = bad[0]
    ev = np.nonzero((f[p].view(np.uint32) != g[p].view(np.uint32)).any(axis=1))[0]
    e = ev[0] if ev.size else 0
    print('packet', p, 'counts', ci[p], cg[p], 'first differing event', e)
    np.set_printoptions(precision=9, linewidth=200)
    for k in range(max(e - 2, 0), e + 2):
        print(k, 'gpu', f[p, k])
        print(k, 'ref', g[p, k])
hits = np.nonzero(g[:, 0, 6] > 0)[0][:3]
for p in hits:
    print('hit packet', p, 'counts', ci[p], cg[p])
    for k in range(0, 3):
        print(k, 'gpu', f[p, k])
        print(k, 'ref', g[p, k])
