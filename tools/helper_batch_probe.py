"""Developer probe: McRunHelper.run_batch streamed through the sweep driver against the
reference's loop of run_one calls (same samples, throughput mode).
usage: python tools/helper_batch_probe.py [samples] [packets]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from test_mcrun import _skin_helper

samples = int(float(sys.argv[1])) if len(sys.argv) > 1 else 64
packets = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
musr = list(np.linspace(5e2, 40e2, samples))
for streamed in (True, False):
    h = _skin_helper(musr*2)
    h.streamed = streamed
    h.run_batch(4, packets)                     # compile + warm-up
    h.sample = -1
    t0 = time.perf_counter()
    batch = h.run_batch(samples, packets)
    dt = time.perf_counter() - t0
    r = np.array([b['detectors']['reflectance'].sum() for b in batch])
    print('%-10s %d samples x %.0e packets: %.1f ms per sample, %.3e packets/s, mean R %.5f'
          % ('streamed' if streamed else 'run_one', samples, packets, 1e3*dt/samples,
             samples*packets/dt, r.mean()))
