"""Developer probe: stage times of the time-resolved configurations (kernel, device-side
trace filter, sampling volume).  usage: python tools/c4_stages.py c4_trace|c4_trace_vox [packets]"""
import importlib
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else benchcfg.PACKETS[name]
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
sim = benchcfg.CONFIGS[name](mc)
make_sv = benchcfg.SAMPLING_VOLUMES[name]
for rep in range(3):
    sim.run(n, download=False)
    sim._stream.synchronize()
    k_ms = sim.run_report['kernel_ms']
    t0 = time.perf_counter()
    sim.filter_trace_on_device(n, download=False)
    sim._stream.synchronize()
    t1 = time.perf_counter()
    sim.sampling_volume(None, make_sv(mc), download=False)
    sim._stream.synchronize()
    t2 = time.perf_counter()
    rr = sim.run_report
    print('%s n=%.0e kernel %.3f ms | filter %.3f ms (wall %.3f) accepted %s | sampling volume %.3f ms (wall %.3f) steps %s' % (
        name, n, k_ms, rr.get('filter_ms', -1), (t1 - t0)*1e3, rr.get('filter_accepted'),
        rr.get('sv_kernel_ms', -1), (t2 - t1)*1e3, rr.get('sv_steps')), flush=True)
