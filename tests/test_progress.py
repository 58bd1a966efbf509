"""ProgressMonitor (mirror of xopto/mcbase/mcprogress.py): control flow on the CPU,
live counter reads during a kernel on the GPU."""
import time

import numpy as np
import pytest

from helpers import build_sim


def test_monitor_control_flow_without_device():
    sim, _, mc = build_sim('mcml_c1_slab')
    calls = []
    mon = mc.mcprogress.ProgressMonitor(sim, interval=0.01, cb=lambda m: calls.append(m.processed()))
    assert mon.start(1000, terminate=False) is mon
    assert mon.target() == 1000 and mon.processed() == 0 and mon.progress() == 0.0
    time.sleep(0.05)                 # no device buffers yet: polls are no-ops
    assert calls == []
    mon.stop()
    assert mon.processed() == 1000 and mon.progress() == 1.0
    mon.resume(2000)
    assert mon.target() == 2000
    mon.terminate()
    with pytest.raises(RuntimeError):
        mon.start(10)


@pytest.mark.gpu
def test_monitor_reads_the_packet_counter_while_the_kernel_runs():
    sim, _, mc = build_sim('mcml_c1_slab')
    sim.run(1000)                    # build + allocate
    n = 60_000_000                   # ~60 ms of kernel
    seen = []
    with mc.mcprogress.ProgressMonitor(
            sim, interval=0.002, cb=lambda m: seen.append((m.processed(), m.threads()))).start(n) as mon:
        _, _, det = sim.run(n)
        time.sleep(0.02)
        final = mon.processed()
    done = np.array([s[0] for s in seen])
    assert len(done) >= 3, seen
    assert (np.diff(done) > 0).all() and done.max() <= n
    assert ((done > 0) & (done < n)).sum() >= 2          # genuinely mid-run reads
    assert final == n or done[-1] >= 0.9*n
    assert det.top.nphotons == n
