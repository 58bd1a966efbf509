"""Developer timing probe (not the judged bench): packets/s of a named case."""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import build_sim

name = sys.argv[1] if len(sys.argv) > 1 else 'mcml_c1_slab'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**7
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sim, geom, mc = build_sim(name)
sim.run(10000, download=False)
for i in range(reps):
    t = time.perf_counter()
    sim.run(n, download=False)
    dt = time.perf_counter() - t
    rr = sim.run_report
    print('{} n={:.0e} kernel {:.2f} ms -> {:.3e} packets/s | wall {:.3f} s | grid {} x {} regs {} smem {} chunk {}'.format(
        name, n, rr['kernel_ms'], n/rr['kernel_ms']*1e3, dt, rr['grid'], rr['block'],
        rr['kernel_attributes']['num_regs'], rr['shared_bytes'], rr['chunk']), flush=True)
