"""Developer timing probe of the C4 pipeline stages: python tools/c4_probe.py [n]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import benchcfg
from pyxopto_b200.mcml import mc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
sim = benchcfg.c4_trace(mc)
sv = benchcfg.c4_sampling_volume(mc)
for rep in range(3):
    t0 = time.perf_counter(); sim.run(n, download=False); sim._stream.synchronize()
    t1 = time.perf_counter(); rr = dict(sim.run_report)
    sim.filter_trace_on_device(n, download=False)
    t2 = time.perf_counter(); fms = sim.run_report['filter_ms'], sim.run_report['filter_accepted']
    sim.sampling_volume(None, sv)
    t3 = time.perf_counter()
    print('run %.1f ms (build %.1f upload %.1f exec %.1f kernel %.2f) | filter %.1f ms (kernels %.2f, accepted %d) | sv %.1f ms (kernel %.2f, steps %d, up %.1f dn %.1f)' % (
        1e3*(t1-t0), 1e3*rr['build'], 1e3*rr['upload'], 1e3*rr['execution'], rr['kernel_ms'],
        1e3*(t2-t1), fms[0], fms[1], 1e3*(t3-t2), sim.run_report['sv_kernel_ms'], sim.run_report['sv_steps'],
        1e3*sim.run_report['upload'], 1e3*sim.run_report['download']), flush=True)
t0 = time.perf_counter(); trace, _, det = sim.run(n); t1 = time.perf_counter()
sv2 = sim.sampling_volume(trace, sv); t2 = time.perf_counter()
print('e2e run %.1f ms (download %.1f) sv %.1f ms; accepted %d dropped %d' % (1e3*(t1-t0), 1e3*sim.run_report.get('download', 0), 1e3*(t2-t1), trace.nphotons, trace.dropped))
