"""Developer probe: where the host time of Mc.run() with host results goes.
usage: python tools/e2e_profile.py c3_vox [packets]"""
import cProfile
import importlib
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else benchcfg.PACKETS[name]
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
sim = benchcfg.CONFIGS[name](mc)


def step():
    trace, fluence, detectors = sim.run(n)
    return float(fluence.raw.sum()) if fluence is not None else 0.0


for _ in range(3):
    step()
t0 = time.perf_counter()
for _ in range(5):
    step()
print('e2e step %.3f ms' % ((time.perf_counter() - t0)*200))
rr = sim.run_report
print({k: (round(v*1e3, 3) if isinstance(v, float) else v) for k, v in rr.items()
       if k in ('upload', 'execution', 'download', 'build', 'kernel_ms')})
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
