"""Developer probe: host time of the C4 step (run + device filter + sampling volume).
usage: python tools/c4_host_profile.py [c4_trace|c4_trace_vox]"""
import cProfile
import importlib
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1] if len(sys.argv) > 1 else 'c4_trace'
n = benchcfg.PACKETS[name]
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
sim = benchcfg.CONFIGS[name](mc)


def step():
    sim.run(n, download=False)
    k = sim.run_report['kernel_ms']
    sim.filter_trace_on_device(n, download=False)
    sim.sampling_volume(None, benchcfg.SAMPLING_VOLUMES[name](mc), download=False)
    return k + sim.run_report['sv_kernel_ms'] + sim.run_report['filter_ms']


for _ in range(5):
    step()
t0 = time.perf_counter()
dev = 0.0
for _ in range(20):
    dev += step()
print('%s step %.3f ms wall, %.3f ms in kernels' % (name, (time.perf_counter() - t0)*50, dev/20))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
