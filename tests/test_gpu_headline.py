"""GPU parity of BASELINE.json's configurations THEMSELVES (benchcfg.py, full
size) - run on the B200 box, through libxopto_b200.so:

  * deterministic mode, bit-exact against the oracle (which the CPU suite pins bit
    for bit to the reference kernel on the same configurations,
    test_oracle_golden.py) - C2's 5-entry stack with the 250 x 500 FluenceRz grid,
    C3 at 201^3 voxels, C4's maxlen-512 trace, the C5 points;
  * throughput mode against the REFERENCE KERNEL ITSELF (oracle/_ref/libref_<config>.so
    on inputs packed by the reference's host layer, oracle/build_ref.py), north
    star's fast-mode criterion: 3 sigma and 1e-3 relative at 1e8 packets (1e-3
    scaled by sqrt(1e8/n) for the run sizes here) on every detector total, every
    detector bin and the fluence marginals along every axis;
  * deterministic-mode Trace trajectories within 1e-5 of the reference kernel's.
"""
import os

import numpy as np
import pytest

import cases
import xo_oracle
from helpers import build_sim, run_size

pytestmark = pytest.mark.gpu

K = float(0x7FFFFF)


def _det_sim(name, **kw):
    from pyxopto_b200.mcbase import mcoptions
    return build_sim(name, options=[mcoptions.McDeterministic.on], **kw)


@pytest.mark.parametrize('name', sorted(cases.BENCH_RUN))
def test_bench_configuration_deterministic_bit_exact(name):
    sim, geom, _ = _det_sim(name)
    n, _ = run_size(name)
    n *= 4
    threads, block = 256, 64
    sim.run(n, maxthreads=threads, wgsize=block, download=False)
    assert sim.run_report['launched_threads'] == threads
    accu, ints, floats = sim.download_raw()
    x_after = sim.download_seeds()[:threads]
    desc = xo_oracle.describe(sim, geom)
    ref = xo_oracle.run(desc, n, threads, sim.rng_seeds_x[:threads],
                        sim.rng_seeds_a[:threads], math=xo_oracle.MATH_PORTABLE)
    assert ref['accu'].sum() > 0
    assert np.array_equal(accu, ref['accu'])
    assert np.array_equal(ints, ref['ints'])
    assert np.array_equal(floats.view(np.uint32), ref['floats'].view(np.uint32))
    assert np.array_equal(x_after, ref['rng_x'][:threads])
    assert sim.run_report['threads'] == ref['num_kernels']


# (packets on the GPU, packets through the reference kernel on the host cores)
THROUGHPUT_RUNS = {
    'c1_slab': (10**7, 4*10**6), 'c2_skin': (10**7, 10**7), 'c3_vox': (10**7, 3*10**6),
    'c5_cyl': (10**7, 4*10**6), 'c5_slab': (10**7, 2*10**6), 'c4_trace': (50000, 50000),
}


def _bound_sigma(p, n_gpu, n_ref):
    # a packet contributes a weight in [0, 1] to a bin / a total: var <= mean
    return np.sqrt(np.maximum(p, 1e-9)*(1.0/n_gpu + 1.0/n_ref))


@pytest.mark.parametrize('config', sorted(THROUGHPUT_RUNS))
def test_throughput_mode_against_the_reference_kernel(config):
    import refbench
    import benchcfg
    if not refbench.available(config):
        pytest.skip('oracle/_ref is not built (needs /root/reference at build time)')
    n_gpu, n_ref = THROUGHPUT_RUNS[config]
    geom = benchcfg.GEOMETRY[config]
    sim, _, _ = build_sim(config)
    sim.device_trace_filter = False
    sim.run(n_gpu, download=False)
    accu, ints, floats = sim.download_raw()
    threads = os.cpu_count() or 8
    ref = refbench.run_reference(config, geom, n_ref, threads, 'ieee')
    assert ref['accu'].size == accu.size
    rel_tol = 1e-3*np.sqrt(1e8/min(n_gpu, n_ref))

    def check(gpu, cpu, scale, what, total):
        g = np.asarray(gpu, np.float64)/scale/n_gpu
        c = np.asarray(cpu, np.float64)/scale/n_ref
        sig = _bound_sigma(c, n_gpu, n_ref)
        if total:
            assert abs(g - c) <= 3*sig + 1e-7, (config, what, float(g), float(c), float(sig))
            if c > 0.01:
                assert abs(g - c) <= rel_tol*c, (config, what, float(g), float(c))
        else:
            bad = np.abs(g - c) > 3*sig + rel_tol*c + 1e-7
            assert not bad.any(), (config, what, np.flatnonzero(bad)[:5], g[bad][:5], c[bad][:5])

    checked = 0
    for det in sim.detectors or ():
        for a in sim.cl_rw_accumulator_allocator.allocations(det):
            g, c = accu[a.offset:a.offset + a.size], ref['accu'][a.offset:a.offset + a.size]
            check(g.sum(dtype=np.float64), c.sum(dtype=np.float64), K, type(det).__name__, True)
            if len(a.shape) > 1:
                gg, cc = g.reshape(a.shape), c.reshape(a.shape)
                for axis in range(gg.ndim):
                    other = tuple(i for i in range(gg.ndim) if i != axis)
                    check(gg.sum(axis=other), cc.sum(axis=other), K,
                          (type(det).__name__, 'axis', axis), False)
            else:
                check(g, c, K, (type(det).__name__, 'bins'), False)
            checked += 1
    flu = sim.fluence
    if flu is not None:
        for a in sim.cl_rw_accumulator_allocator.allocations(flu):
            shape = a.shape
            if type(flu).__name__ == 'FluenceRz':
                shape = (flu.shape[1], flu.shape[0])         # bins are z-major
            g = accu[a.offset:a.offset + a.size].reshape(shape)
            c = ref['accu'][a.offset:a.offset + a.size].reshape(shape)
            k = float(flu.k)
            check(g.sum(dtype=np.float64), c.sum(dtype=np.float64), k, 'fluence total', True)
            for axis in range(g.ndim):
                other = tuple(i for i in range(g.ndim) if i != axis)
                check(g.sum(axis=other, dtype=np.float64), c.sum(axis=other, dtype=np.float64),
                      k, ('fluence marginal', axis), False)
            checked += 1
    assert checked > 0
    if sim.trace is not None:
        # per-packet event counts and terminal events of the trace rows
        tp = sim._packed['trace']
        ml = int(sim.trace.maxlen)
        co, do = int(tp.count_buffer_offset), int(tp.data_buffer_offset)

        def unpack(ints_, floats_, n):
            cnt = ints_[co:co + n].astype(np.float64)
            rows = floats_[do:do + n*ml*8].reshape(n, ml, 8)
            last = np.minimum(cnt.astype(np.int64), ml) - 1
            return cnt, rows[np.arange(n), np.maximum(last, 0)]

        cg, tg = unpack(ints, floats, n_gpu)
        cr, tr = unpack(ref['ints'], ref['floats'], n_ref)

        def close(a, b, what):
            se = np.sqrt(a.var()/a.size + b.var()/b.size)
            assert abs(a.mean() - b.mean()) <= 4*se + 1e-12, (what, a.mean(), b.mean(), se)

        close(cg, cr, 'events per packet')
        close((cg >= ml).astype(float), (cr >= ml).astype(float), 'overflow fraction')
        ok_g, ok_r = cg < ml, cr < ml
        for col, what in ((0, 'x'), (1, 'y'), (2, 'z'), (5, 'pz'), (6, 'w'), (7, 'pl')):
            close(tg[ok_g, col].astype(np.float64), tr[ok_r, col].astype(np.float64),
                  'terminal ' + what)


@pytest.mark.parametrize('name', sorted(cases.TRAJ_RUN))
def test_deterministic_trajectories_within_1e5_of_the_reference_kernel(name):
    """GPU deterministic-mode trace rows next to the reference kernel's own rows
    (tests/golden/traj_<name>.npz, one packet per work-item)."""
    from test_oracle_golden import check_trajectories
    sim, geom, _ = _det_sim(name)
    sim.device_trace_filter = False
    n = cases.TRAJ_RUN[name]
    sim.run(n, maxthreads=n, wgsize=64, download=False)
    assert sim.run_report['launched_threads'] == n
    accu, ints, floats = sim.download_raw()
    tp = sim._packed['trace']
    ml = int(sim.trace.maxlen)
    co, do = int(tp.count_buffer_offset), int(tp.data_buffer_offset)
    check_trajectories(name, floats[do:do + n*ml*8].reshape(n, ml, 8), ints[co:co + n],
                       accu, sim.download_seeds()[:n])
