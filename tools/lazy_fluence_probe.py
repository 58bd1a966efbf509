"""Developer probe: run(out=...) in batches with the fluence result on the host (reference flow)
and on the device (Mc.lazy_fluence).  usage: python tools/lazy_fluence_probe.py [config] [packets] [batches]"""
import importlib
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1] if len(sys.argv) > 1 else 'c3_vox'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000000
batches = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
for lazy in (False, True):
    sim = benchcfg.CONFIGS[name](mc)
    sim.lazy_fluence = lazy
    for rep in range(3):
        t0 = time.perf_counter()
        out = None
        for _ in range(batches):
            out = sim.run(n, out=out)
        total = float(out[1].raw.sum())
        dt = time.perf_counter() - t0
    print('%s lazy_fluence=%d: %d x %.0e packets in %.1f ms -> %.4e packets/s (kernel %.2f ms per batch, sum %.6e)' % (
        name, lazy, batches, n, dt*1e3, batches*n/dt, sim.run_report['kernel_ms'], total))
