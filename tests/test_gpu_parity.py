"""GPU parity tests (run on the B200 box): the CUDA path, called through
libxopto_b200.so, against the CPU oracle on the same seeded inputs.

Deterministic mode (static block schedule, IEEE arithmetic, portable elementary
functions) must be BIT-EXACT: 64-bit fixed-point accumulators, trace buffers
(float bit patterns), per-packet event counts and the advanced MWC states.
Throughput mode (dynamic schedule, MUFU math) must agree statistically.
"""
import numpy as np
import pytest

import cases
import xo_oracle
from helpers import build_sim, golden, run_size

pytestmark = pytest.mark.gpu


def _det_sim(name, **kw):
    from pyxopto_b200.mcbase import mcoptions
    return build_sim(name, options=[mcoptions.McDeterministic.on], **kw)


def test_library_sees_a_b200():
    from pyxopto_b200.cu import abi
    assert abi.device_count() >= 1
    info = abi.device_info(0)
    assert info['cc_major'] >= 9, info


def test_rng_stream_matches_reference_mapping():
    sim, _, _ = build_sim('mcml_c1_slab')
    for t in (0, 1, 77):
        x, a = int(sim.rng_seeds_x[t]), int(sim.rng_seeds_a[t])
        dev = sim.rng_test(64, x, a)
        assert np.array_equal(dev, xo_oracle.rng_test(x, a, 64))


@pytest.mark.parametrize('fn', ['log', 'sincos', 'cbrt', 'pow', 'exp', 'atan2', 'sqrt', 'div'])
def test_deterministic_math_bit_exact(fn):
    from pyxopto_b200.mcbase import rngkernel
    sim, _, _ = build_sim('mcml_c1_slab')
    sim._ensure_device()
    rs = np.random.RandomState(11)
    n = 200000
    u = rs.rand(n).astype(np.float32)
    v = rs.rand(n).astype(np.float32)
    if fn == 'log':
        a, b = np.concatenate([u, [0.0, 1.0, 2.0**-32, 3.4e38]]).astype(np.float32), None
    elif fn == 'sincos':
        a, b = (u*np.float32(6.2831855)).astype(np.float32), None
    elif fn == 'cbrt':
        a, b = (2*u - 1).astype(np.float32), None
    elif fn == 'pow':
        a, b = (u + np.float32(0.01)).astype(np.float32), (-4*v).astype(np.float32)
    elif fn == 'exp':
        a, b = (-20*u).astype(np.float32), None
    elif fn == 'atan2':
        a, b = (2*u - 1).astype(np.float32), (2*v - 1).astype(np.float32)
    else:
        a, b = (u + np.float32(1e-3)).astype(np.float32), (v + np.float32(1e-3)).astype(np.float32)
    d0, d1 = rngkernel.math_probe(sim, fn, a, b, deterministic=True)
    o0, o1 = xo_oracle.math_probe(fn, xo_oracle.MATH_PORTABLE, a, b)
    assert np.array_equal(d0.view(np.uint32), o0.view(np.uint32))
    if fn == 'sincos':
        assert np.array_equal(d1.view(np.uint32), o1.view(np.uint32))


@pytest.mark.parametrize('name', sorted(cases.ALL_CASES) + sorted(cases.UNPINNED_CASES))
def test_deterministic_mode_bit_exact(name):
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
    assert np.array_equal(accu, ref['accu'])
    assert np.array_equal(ints, ref['ints'])
    assert np.array_equal(floats.view(np.uint32), ref['floats'].view(np.uint32))
    assert np.array_equal(x_after, ref['rng_x'][:threads])
    assert sim.run_report['threads'] == ref['num_kernels']


@pytest.mark.parametrize('name', ['mcml_c1_slab', 'mcml_mhg_gauss_cart_flurz',
                                  'mcml_gk_fiber_six_flu', 'mcml_surface_six_lambert',
                                  'mcml_surface_lambert_top', 'mcvox_gauss_fluence',
                                  'mccyl_hg_line_fiz', 'mccyl_mhg_gauss_total_flurz',
                                  'mccyl_hg_isopoint_outside', 'mcvox_isovoxel_fluence',
                                  'mcml_mhg_lambertianfiber_radial', 'mcml_hg_fiber_lineararray',
                                  'mcml_mhg_line_fiberarray', 'mcml_hg_fiber_arrays_pl',
                                  'mcml_hg_fiber_fluencecylt', 'mcml_surface_lineararray',
                                  'mcml_surface_fiberarray', 'mcml_hg_line_totallut',
                                  'mcml_lut_ufiberlut_totallut', 'mcml_hg_line_fiberlutarray',
                                  'mcml_mhg_rect_uniform', 'mcml_hg_rect_lambertian_inside',
                                  'mcvox_isovoxels_total', 'mcml_hgdir_line_radial',
                                  'mcml_hg_ufiberni_radial', 'mcml_mhg_lfiberni_cart',
                                  'mcml_hg_ufiberlutni_total', 'mcml_hg_rectlut_inside',
                                  'mcvox_lfiber_radial', 'mcvox_ufiberlut_fluence',
                                  'mcml_aniso_line_cart_flu', 'mcml_rayleigh_line_radial'])
def test_throughput_mode_statistics(name):
    """Fast mode vs oracle (libm, different schedule): totals within 4 sigma."""
    sim, geom, _ = build_sim(name)
    n = 200000
    sim.run(n, download=False)
    accu, _, _ = sim.download_raw()
    desc = xo_oracle.describe(sim, geom)
    ref = xo_oracle.run(desc, n, 64, sim.rng_seeds_x[:64], sim.rng_seeds_a[:64],
                        math=xo_oracle.MATH_LIBM)
    K = 0x7FFFFF
    owners = [d for d in sim.detectors] + ([sim.fluence] if sim.fluence else [])
    for owner in owners:
        for a in sim.cl_rw_accumulator_allocator.allocations(owner):
            tot_gpu = accu[a.offset:a.offset + a.size].sum()/K/n
            tot_ref = ref['accu'][a.offset:a.offset + a.size].sum()/K/n
            # per-packet weight is in [0,1]: sigma of the mean <= sqrt(p/n)
            sigma = np.sqrt(max(tot_ref, 1e-6)/n)*np.sqrt(2)
            assert abs(tot_gpu - tot_ref) <= 4*sigma + 1e-5, (
                type(owner).__name__, tot_gpu, tot_ref)


@pytest.mark.parametrize('name', ['mcvox_gauss_fluence', 'mcvox_isopoint_fluencerate',
                                  'mcml_mhg_gauss_cart_flurz', 'mcml_surface_six_lambert'])
def test_throughput_mode_profiles(name):
    """Fast mode vs oracle beyond totals: every bin of the marginal profiles of
    the fluence grid (along each axis) and every detector bin must agree within
    5 sigma.  The throughput loops re-associate the geometry (mcvox: one ray per
    flight + incremental voxel walk), so this pins where the energy goes."""
    sim, geom, _ = build_sim(name)
    n = 400000
    sim.run(n, download=False)
    accu, _, _ = sim.download_raw()
    desc = xo_oracle.describe(sim, geom)
    ref = xo_oracle.run(desc, n, 64, sim.rng_seeds_x[:64], sim.rng_seeds_a[:64],
                        math=xo_oracle.MATH_LIBM)
    K = 0x7FFFFF

    def check(gpu, cpu, scale, what):
        gpu = gpu.astype(np.float64)/scale/n
        cpu = cpu.astype(np.float64)/scale/n
        # per-packet contribution to a bin is in [0, 1] (weight units)
        sigma = np.sqrt(np.maximum(cpu, 1e-6)/n)*np.sqrt(2)
        bad = np.abs(gpu - cpu) > 5*sigma + 2e-5
        assert not bad.any(), (what, np.flatnonzero(bad)[:5], gpu[bad][:5], cpu[bad][:5])

    for det in sim.detectors or ():
        for a in sim.cl_rw_accumulator_allocator.allocations(det):
            check(accu[a.offset:a.offset + a.size], ref['accu'][a.offset:a.offset + a.size],
                  K, type(det).__name__)
    flu = sim.fluence
    for a in sim.cl_rw_accumulator_allocator.allocations(flu):
        g = accu[a.offset:a.offset + a.size].reshape(a.shape)
        c = ref['accu'][a.offset:a.offset + a.size].reshape(a.shape)
        k = float(flu.k)
        if sim.resolved_options().get('MC_FLUENCE_MODE_RATE'):
            # rate mode stores weight/mua: normalise to weight units with the
            # smallest mua so the [0, 1] bound on a contribution holds
            k *= 1.0/min(m.mua for m in sim.materials if m.mua > 0) if geom == 'mcvox' else 1.0
        for axis in range(g.ndim):
            other = tuple(i for i in range(g.ndim) if i != axis)
            check(g.sum(axis=other), c.sum(axis=other), k, ('fluence', axis))


@pytest.mark.parametrize('config,n', [('c1_slab', 10**8), ('c2_skin', 10**8),
                                      ('c3_vox', 10**8), ('c5_cyl', 2*10**7)])
def test_fast_mode_agrees_with_deterministic_mode_at_scale(config, n):
    """BASELINE.json's fast-mode criterion at full size: totals of every detector
    and of the fluence grid from the throughput kernel agree with the
    deterministic kernel (bit-exact against the oracle, see above) within 3 sigma
    and 1e-3 relative at 1e8 packets (1e-3 scaled by sqrt(1e8/n) for the smaller
    runs); the fluence depth profile agrees bin by bin within 4 sigma."""
    import importlib
    import benchcfg
    from pyxopto_b200.mcbase import mcoptions
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(benchcfg.GEOMETRY[config]))
    res = []
    for opts in ([], [mcoptions.McDeterministic.on]):
        sim = benchcfg.CONFIGS[config](mc, options=opts)
        sim.run(n, download=False)
        res.append((sim, sim.download_raw()[0]))
    (sim, fast), (_, det) = res
    rel_tol = 1e-3*np.sqrt(1e8/n)
    K = float(0x7FFFFF)
    owners = [d for d in (sim.detectors or ())] + ([sim.fluence] if sim.fluence else [])
    checked = 0
    for owner in owners:
        scale = float(owner.k) if owner is sim.fluence else K
        for a in sim.cl_rw_accumulator_allocator.allocations(owner):
            f = fast[a.offset:a.offset + a.size].astype(np.float64)/scale/n
            d = det[a.offset:a.offset + a.size].astype(np.float64)/scale/n
            tf, td = f.sum(), d.sum()
            sigma = np.sqrt(max(td, 1e-9)/n)*np.sqrt(2)
            assert abs(tf - td) <= 3*sigma + 1e-7, (type(owner).__name__, tf, td, sigma)
            if td > 0.01:
                assert abs(tf - td) <= rel_tol*td, (type(owner).__name__, tf, td)
            checked += 1
            if owner is sim.fluence:
                # depth profile: sum over all but the slowest (z) axis
                fz = f.reshape(a.shape).reshape(a.shape[0], -1).sum(axis=1) if len(a.shape) > 1 \
                    else f
                dz = d.reshape(a.shape).reshape(a.shape[0], -1).sum(axis=1) if len(a.shape) > 1 \
                    else d
                if type(owner).__name__ == 'FluenceRz':
                    fz = f.reshape(-1, owner.shape[0]).sum(axis=1)
                    dz = d.reshape(-1, owner.shape[0]).sum(axis=1)
                sig = np.sqrt(np.maximum(dz, 1e-9)/n)*np.sqrt(2)
                bad = np.abs(fz - dz) > 4*sig + 1e-7
                assert not bad.any(), (np.flatnonzero(bad)[:5], fz[bad][:5], dz[bad][:5])
    assert checked > 0


@pytest.mark.parametrize('name', ['mcvox_line_mhg_trace', 'mcvox_line_mhg_trace:lane-resident',
                                  'mcvox_line_mhg_trace_startend',
                                  'mcvox_line_mhg_trace_startend:lane-resident',
                                  'mcml_lut_iso_radialpl_trace', 'mccyl_gk_ubeam_fiz_trace'])
def test_throughput_mode_trace_statistics(name):
    """Trace recording in the throughput loops (mcvox: one event per crossing /
    interaction of the voxel walk, like the reference's loop trips): the
    distribution of per-packet event counts, the overflow fraction and the
    terminal events (position, weight, path length) agree with the oracle."""
    n = 40000
    name, _, loop = name.partition(':')
    sim, geom, _ = build_sim(name)
    if geom == 'mcvox':
        # the packet pool records traces as well; by default a full trace keeps the
        # lane-resident loop (store-bound, measured faster there)
        sim.pool_full_trace = not loop
        sim.pool_slots = 0 if loop else 64
    sim.run(n, download=False)
    if geom == 'mcvox':
        assert sim.run_report['loop'] == ('lane-resident rays' if loop else 'packet pool')
    accu, ints, floats = sim.download_raw()
    desc = xo_oracle.describe(sim, geom)
    ref = xo_oracle.run(desc, n, 64, sim.rng_seeds_x[:64], sim.rng_seeds_a[:64],
                        math=xo_oracle.MATH_LIBM)
    tp = sim._packed['trace']
    maxlen = int(sim.trace.maxlen)
    co, do = int(tp.count_buffer_offset), int(tp.data_buffer_offset)

    def unpack(ints_, floats_):
        cnt = ints_[co:co + n].astype(np.float64)
        rows = floats_[do:do + n*maxlen*8].reshape(n, maxlen, 8)
        last = np.minimum(cnt.astype(np.int64), maxlen) - 1
        term = rows[np.arange(n), np.maximum(last, 0)]
        return cnt, term

    cg, tg = unpack(ints, floats)
    cr, tr = unpack(ref['ints'], ref['floats'])
    assert cg.min() >= 1

    def close(a, b, what, k=5.0):
        # (standard error of the difference of the two sample means: the terminal
        # events are compared on the sub-samples that did not overflow)
        se = np.sqrt(a.var()/max(a.size, 1) + b.var()/max(b.size, 1))
        assert abs(a.mean() - b.mean()) <= k*se + 1e-9, (what, a.mean(), b.mean(), se)

    close(cg, cr, 'events per packet')
    close((cg > maxlen).astype(float), (cr > maxlen).astype(float), 'overflow fraction')
    # (a start / end trace has maxlen = 2 = its event count, mctrace.py:848-854)
    ok_g, ok_r = cg <= maxlen, cr <= maxlen
    assert ok_g.any() and ok_r.any()
    for col, what in ((0, 'x'), (1, 'y'), (2, 'z'), (5, 'pz'), (6, 'w'), (7, 'pl')):
        close(tg[ok_g, col].astype(np.float64), tr[ok_r, col].astype(np.float64),
              'terminal ' + what)
    # first recorded event is the launch in both
    fg = floats[do:do + n*maxlen*8].reshape(n, maxlen, 8)[:, 0, :]
    fr = ref['floats'][do:do + n*maxlen*8].reshape(n, maxlen, 8)[:, 0, :]
    close(fg[:, 6].astype(np.float64), fr[:, 6].astype(np.float64), 'launch weight')


@pytest.mark.parametrize('maxlen', [512, 64, 8])
def test_full_trace_rows_are_consistent(maxlen):
    """Config 4's trace stream leaves the throughput kernel as one 256-bit store
    (one 32-byte event = one DRAM sector) per lane and loop trip.  Every packet's
    row must be a coherent trajectory (launch event first, unit directions, z
    inside the slab, optical path length non-decreasing, terminal event last,
    zero tail) and the per-packet statistics must agree with the oracle."""
    import benchcfg
    from pyxopto_b200.mcml import mc
    n = 30000
    sim = benchcfg.c4_trace(mc, maxlen=maxlen)
    sim.device_trace_filter = False          # raw rows wanted: zero-filled tails
    assert sim._pack(n) is not None and sim._trace_aligned() == 2
    assert 'XO_TRACE_ALIGNED 2' in sim.kernel_source()
    sim.run(n, download=False)
    accu, ints, floats = sim.download_raw()
    tp = sim._packed['trace']
    co, do = int(tp.count_buffer_offset), int(tp.data_buffer_offset)
    cnt = ints[co:co + n]
    rows = floats[do:do + n*maxlen*8].reshape(n, maxlen, 8)
    assert cnt.min() >= 2
    nrec = np.minimum(cnt, maxlen)
    idx = np.arange(maxlen)[None, :]
    valid = idx < nrec[:, None]
    # zero tails, non-zero recorded events (|dir| = 1)
    assert not rows[~valid].any()
    dn = np.linalg.norm(rows[..., 3:6], axis=2)
    assert np.allclose(dn[valid], 1.0, atol=1e-4)
    # launch event first: Line source at the origin, weight 1 - Rspecular
    assert np.allclose(rows[:, 0, :3], 0.0) and np.allclose(rows[:, 0, 5], 1.0)
    assert np.allclose(rows[:, 0, 6], 1.0 - ((1.33 - 1)/(1.33 + 1))**2, atol=1e-6)
    z = rows[..., 2]
    assert z[valid].min() >= -1e-9 and z[valid].max() <= 10e-3 + 1e-8
    # optical path length never decreases along a row (rows are not interleaved);
    # the last slot of an overflowed row was overwritten by later events
    pl = rows[..., 7]
    ok = cnt <= maxlen
    d = np.diff(pl, axis=1)
    pair_valid = valid[:, 1:] & ok[:, None]
    assert (d[pair_valid] >= 0).all()
    # packets that did not overflow end on a surface or by the lottery / rmax
    last = rows[np.arange(n), np.maximum(nrec - 1, 0)]
    # statistics against the oracle
    desc = xo_oracle.describe(sim, 'mcml')
    ref = xo_oracle.run(desc, n, 64, sim.rng_seeds_x[:64], sim.rng_seeds_a[:64],
                        math=xo_oracle.MATH_LIBM)
    rcnt = ref['ints'][co:co + n]
    rrows = ref['floats'][do:do + n*maxlen*8].reshape(n, maxlen, 8)
    rlast = rrows[np.arange(n), np.maximum(np.minimum(rcnt, maxlen) - 1, 0)]

    def close(a, b, what):
        a, b = a.astype(np.float64), b.astype(np.float64)
        se = np.sqrt(a.var()/max(a.size, 1) + b.var()/max(b.size, 1))
        assert abs(a.mean() - b.mean()) <= 5*se + 1e-12, (what, a.mean(), b.mean(), se)

    close(cnt, rcnt, 'events per packet')
    close((cnt >= maxlen), (rcnt >= maxlen), 'overflow fraction')
    for col, what in ((2, 'z'), (6, 'w'), (7, 'pl')):
        close(last[ok, col], rlast[rcnt <= maxlen, col], 'terminal ' + what)
    # interior events too: mean depth of the k-th event
    for k in (1, 2, 3, 4, 5, 7):
        if k < maxlen - 1:
            close(rows[cnt > k, k, 2], rrows[rcnt > k, k, 2], 'z of event %d' % k)


def test_run_returns_reference_style_results():
    sim, _, mc = build_sim('mcml_c1_slab')
    trace, fluence, detectors = sim.run(100000)
    assert trace is None and fluence is None
    r = detectors.top.reflectance
    assert r.shape == (1000,)
    total = detectors.top.raw.sum()/detectors.top.nphotons
    spec = detectors.specular.raw.sum()/detectors.specular.nphotons
    assert abs(spec - ((1.33 - 1)/(1.33 + 1))**2) < 1e-6     # Fresnel, SURVEY 8c
    assert 0.36 < total < 0.42
    # continuation: accumulate into previous results
    _, _, detectors2 = sim.run(100000, out=(None, None, detectors))
    assert detectors2.top.nphotons == 200000


@pytest.mark.parametrize('name', cases.SV_CASES)
def test_sampling_volume_bit_exact_and_fast(name):
    """Mc.sampling_volume on the device vs the oracle's restatement of the
    reference SamplingVolume kernel, fed with the same trace rows: deterministic
    mode bit-exact (64-bit voxel accumulators + total weight); throughput mode
    (MUFU division / square root) within 1e-4 of the grid total."""
    sim, geom, mc = _det_sim(name)
    n = run_size(name)[0]
    trace, _, _ = sim.run(n, maxthreads=256, wgsize=64)
    assert trace.nphotons == n
    sv = cases.make_sv(mc, name)
    sim.sampling_volume(trace, sv)
    accu, _, _ = sim.download_raw()
    tp, sp = sim._packed['sv_trace'], sim._packed['sv']
    ints = np.zeros(max(sim.cl_rw_int_allocator.size, 1), np.int32)
    floats = np.zeros(max(sim.cl_rw_float_allocator.size, 1), np.float32)
    ints[tp.count_buffer_offset:tp.count_buffer_offset + n] = trace.n
    rows = np.ascontiguousarray(trace.data).view(np.float32).reshape(-1)
    floats[tp.data_buffer_offset:tp.data_buffer_offset + rows.size] = rows
    ref = xo_oracle.sampling_volume(tp, sp, n, ints, floats,
                                    sim.cl_rw_accumulator_allocator.size)
    assert np.array_equal(accu, ref['accu'])
    assert sv.weight == ref['total_weight']/sv.k
    assert sim.run_report['sv_steps'] == ref['steps']
    assert ref['accu'].sum() > 0
    # reference-style result object
    assert sv.data.shape == sv.shape
    total = float(ref['accu'].sum())
    assert abs(sv.data.sum()*sv.k*sv._multiplier() - total) <= 1e-9*total

    fast, _, mc2 = build_sim(name)
    sv2 = cases.make_sv(mc2, name)
    fast.sampling_volume(trace, sv2)
    assert abs(sv2.data.sum() - sv.data.sum()) <= 1e-4*sv.data.sum()
    assert sv2.weight == pytest.approx(sv.weight, rel=1e-6)
    # the throughput kernel walks the segments of a packet side by side (one warp per
    # packet): voxel by voxel it differs from the reference-structured kernel by the
    # rounding of the individual deposits only
    assert np.allclose(sv2.data, sv.data, rtol=2e-3, atol=2e-5*sv.data.max())
    # ... and the reference-structured kernel in throughput math stays available
    fast.sv_warp_per_packet = False
    sv3 = cases.make_sv(mc2, name)
    fast.sampling_volume(trace, sv3)
    assert abs(sv3.data.sum() - sv.data.sum()) <= 1e-4*sv.data.sum()
    assert np.allclose(sv3.data, sv.data, rtol=2e-3, atol=2e-5*sv.data.max())


def _filters(mc):
    F = mc.mctrace.Filter
    inf = float('inf')
    return {
        'mcml_lut_iso_radialpl_trace': F(
            z=(-inf, 1e-9), pz=(-1.0, -0.3), r=(0.0, 2e-3, (0.1e-3, 0.0)),
            x=[(-1.0, -0.05e-3), (0.05e-3, 1.0)], pl=(0.0, 0.01)),
        'mcvox_line_mhg_trace': F(
            z=[(-inf, 0.0), (0.45e-3, inf)], dir=(0.5, 1.0, (0.0, 0.0, -1.0))),
        'mccyl_gk_ubeam_fiz_trace': F(
            y=(-2e-3, 2e-3), dir=[(0.2, 1.0, (1.0, 0.0, 0.0)), (0.2, 1.0, (-1.0, 0.0, 0.0))]),
    }


@pytest.mark.parametrize('name', cases.SV_CASES)
def test_device_trace_filter_equals_host_filter(name):
    """Filter + stable compaction on the device (8f-1) vs the reference's flow
    (download everything, numpy filter on the host): same rows in the same
    order, same counts, same n_dropped; the compact rows feed sampling_volume
    without leaving the device."""
    from pyxopto_b200.mcbase import mcoptions
    n = run_size(name)[0]*4
    results = []
    for on_device in (True, False):
        sim, geom, mc = build_sim(name, options=[mcoptions.McDeterministic.on])
        sim.trace.filter = _filters(mc)[name]
        sim.device_trace_filter = on_device
        trace, _, _ = sim.run(n, maxthreads=256, wgsize=64)
        results.append((sim, mc, trace))
    (sim_d, mc_d, tr_d), (sim_h, mc_h, tr_h) = results
    assert 0 < tr_h.nphotons < n
    assert tr_d.nphotons == tr_h.nphotons
    assert tr_d.dropped == tr_h.dropped
    assert np.array_equal(tr_d.n, tr_h.n)
    assert tr_d.data.tobytes() == tr_h.data.tobytes()
    # sampling volume from device-resident rows == from re-uploaded host rows
    sv_d = sim_d.sampling_volume(tr_d, cases.make_sv(mc_d, name))
    assert sim_d.run_report['sv_rows_resident'] is True
    sv_h = sim_h.sampling_volume(tr_h, cases.make_sv(mc_h, name))
    assert sim_h.run_report['sv_rows_resident'] is False
    assert np.array_equal(sv_d.data, sv_h.data) and sv_d.weight == sv_h.weight
    assert sv_d.data.sum() > 0


@pytest.mark.parametrize('config', ['c4_trace', 'c4_trace_vox'])
def test_sampling_volume_accumulates_on_the_device(config):
    """``Mc.lazy_sampling_volume``: one SamplingVolume fed by several batches keeps its
    integer grid on the device and comes to the host when ``data`` is read.  The
    result equals the reference flow (one conversion + download per call, summed in
    float64 on the host) to rounding, the total weight exactly; the grid restarts at
    zero afterwards."""
    import importlib
    import benchcfg
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(benchcfg.GEOMETRY[config]))
    from pyxopto_b200.mcbase import mcoptions
    n = 20000 if config == 'c4_trace' else 8000
    results = []
    for lazy in (True, False):
        # (deterministic mode: both simulators trace exactly the same packets)
        sim = benchcfg.CONFIGS[config](mc, maxlen=128, options=[mcoptions.McDeterministic.on])
        sim.lazy_sampling_volume = lazy
        sv = benchcfg.SAMPLING_VOLUMES[config](mc)
        for _ in range(3):
            trace, _, det = sim.run(n, maxthreads=2048, wgsize=64)
            assert trace.nphotons > 0
            out = sim.sampling_volume(trace, sv)
            assert out is sv
            assert (sv._pending is not None) == lazy
        results.append((sim, sv, sv.weight, np.array(sv.data, copy=True)))
    (sim_l, sv_l, w_l, d_l), (sim_e, sv_e, w_e, d_e) = results
    assert sv_l._pending is None and sim_l._sv_resident is None
    assert w_l == w_e and d_l.sum() > 0
    assert np.allclose(d_l, d_e, rtol=1e-12, atol=0.0)
    # a second study with the same object continues from the collected data
    trace, _, _ = sim_l.run(n, maxthreads=2048, wgsize=64)
    sim_l.sampling_volume(trace, sv_l)
    assert sv_l.data.sum() > d_l.sum()


@pytest.mark.parametrize('name', ['mcvox_gauss_fluence', 'mcml_hg_gauss_fluencerzt', 'c3_vox'])
def test_fluence_result_accumulates_on_the_device(name):
    """``Mc.lazy_fluence``: a fluence result fed by several ``run(out=...)`` calls keeps its
    float64 grid on the device (AccuScaleAdd) and comes to the host when ``raw`` is read -
    bit-identical to the reference flow (one download per run, ``raw += accumulators/k`` on
    the host, fluence.py:349-376).  A result that already holds host data continues on the
    device; two results alternate through one simulator."""
    n = 20000
    grids = []
    for lazy in (True, False):
        # (deterministic mode: both simulators run exactly the same packets)
        sim = _det_sim(name)[0]
        sim.lazy_fluence = lazy
        out = None
        for _ in range(3):
            out = sim.run(n, out=out, maxthreads=2048, wgsize=64)
            assert (out[1]._pending is not None) == lazy
        flu = out[1]
        assert flu.nphotons == 3*n
        raw = np.array(flu.raw, copy=True)
        assert flu._pending is None and sim._flu_resident is None
        # the collected result (host data now) goes back to the device for two more runs
        out = sim.run(n, out=out, maxthreads=2048, wgsize=64)
        other = sim.run(n, maxthreads=2048, wgsize=64)          # a second, fresh result
        assert (other[1]._pending is not None) == lazy and out[1]._pending is None
        out = sim.run(n, out=out, maxthreads=2048, wgsize=64)
        assert other[1]._pending is None
        grids.append((raw, np.array(out[1].raw, copy=True), np.array(other[1].raw, copy=True),
                      out[1].nphotons, other[1].nphotons))
    (raw_l, five_l, one_l, n5_l, n1_l), (raw_e, five_e, one_e, n5_e, n1_e) = grids
    assert raw_l.sum() > 0 and raw_l.shape == raw_e.shape
    assert raw_l.tobytes() == raw_e.tobytes()
    assert five_l.tobytes() == five_e.tobytes() and one_l.tobytes() == one_e.tobytes()
    assert (n5_l, n1_l) == (n5_e, n1_e) == (5*n, n)


# ---------------------------------------------------------------------------
# user-written plugins: OpenCL-C fragments compiled through xo_clcompat*.cuh
def _raw_run(name, n, threads=256, block=64, deterministic=True):
    sim = (_det_sim(name)[0] if deterministic else build_sim(name)[0])
    kw = dict(maxthreads=threads, wgsize=block) if deterministic else {}
    sim.run(n, download=False, **kw)
    accu, ints, floats = sim.download_raw()
    sim._raw_trace = (ints, floats)
    return sim, accu, sim.download_seeds()[:threads]


@pytest.mark.parametrize('user, native', [('mcml_user_plugins', 'mcml_user_plugins_native'),
                                          ('mcml_user_surface_reflector', 'mcml_surface_lambert_top'),
                                          ('mcml_user_trace', 'mcml_lut_iso_radialpl_trace'),
                                          ('mcvox_user_trace', 'mcvox_line_mhg_trace'),
                                          ('mccyl_user_trace', 'mccyl_gk_ubeam_fiz_trace')])
def test_user_fragments_equal_builtins_bit_exact(user, native):
    """A user-written phase function, source and detector (tests/user_plugins.py)
    that restate Hg / Line / Radial - and a user-written surface layout that restates
    LambertianReflector: deterministic mode must give the very accumulators and MWC
    states of the built-in CUDA plugins (which the oracle pins,
    test_deterministic_mode_bit_exact[mcml_user_plugins_native / mcml_surface_lambert_top])."""
    n = 4*run_size(user)[0]
    sim_u, accu_u, x_u = _raw_run(user, n)
    sim_n, accu_n, x_n = _raw_run(native, n)
    assert 'xo_clcompat.cuh' in sim_u._last_src and 'xo_clcompat' not in sim_n._last_src
    assert accu_u.sum() > 0
    assert np.array_equal(accu_u, accu_n)
    assert np.array_equal(x_u, x_n)
    assert sim_u.run_report['threads'] == sim_n.run_report['threads']
    # (trace rows and event counts: a user-written trace writes what the built-in one does)
    assert np.array_equal(sim_u._raw_trace[0], sim_n._raw_trace[0])
    assert np.array_equal(sim_u._raw_trace[1].view(np.uint32), sim_n._raw_trace[1].view(np.uint32))


@pytest.mark.parametrize('name', sorted(cases.USER_CASES))
def test_user_fragments_match_reference_kernel(name):
    """The same fragments executed by the reference's own kernel (golden vectors,
    libm math, static schedule): deterministic mode here differs at ulp level only
    (same criterion as test_portable_math_is_statistically_the_reference), and the
    throughput mode agrees within 4 sigma per detector."""
    g = golden(name)
    n, t = run_size(name)
    sim, accu, _ = _raw_run(name, n, threads=t, block=t)
    a, b = float(accu.sum()), float(g['accu'].sum())
    assert abs(a - b) <= max(1e-3*b, 2*(n/t)**0.5*0x7FFFFF)
    same = np.count_nonzero(accu == g['accu'])/g['accu'].size
    # (one work-item that takes another branch after a 1-ulp difference of `log` moves its
    # ~200 packets: a few bins of a detector, a larger share of a 13 440-cell fluence grid)
    assert same > (0.9 if sim.fluence is None or name.startswith('mcml') else 0.6), same
    if sim.trace is not None:
        # a user-written trace: its rows against the rows the reference kernel wrote with
        # the same fragment (north star: 1e-5 relative for the packets that agree in count)
        from helpers import trajectory_agreement
        P = sim._packed['trace']
        do, co, ml = int(P.data_buffer_offset), int(P.count_buffer_offset), int(sim.trace.maxlen)
        ints, floats = sim._raw_trace
        frac, worst, _ = trajectory_agreement(
            floats[do:do + n*ml*8].reshape(n, ml, 8), ints[co:co + n],
            g['floats'][do:do + n*ml*8].reshape(n, ml, 8), g['ints'][co:co + n])
        assert frac > 0.9 and worst <= 1e-5, (frac, worst)
    # throughput mode, more packets, against the golden run scaled
    n_fast = 200000
    sim_f, accu_f, _ = _raw_run(name, n_fast, deterministic=False)
    K = 0x7FFFFF
    for det in list(sim_f.detectors) + ([sim_f.fluence] if sim_f.fluence is not None else []):
        for al in sim_f.cl_rw_accumulator_allocator.allocations(det):
            tot_gpu = accu_f[al.offset:al.offset + al.size].sum()/K/n_fast
            tot_ref = g['accu'][al.offset:al.offset + al.size].sum()/K/n
            sigma = np.sqrt(max(tot_ref, 1e-6)*(1.0/n + 1.0/n_fast))*np.sqrt(2)
            assert abs(tot_gpu - tot_ref) <= 4*sigma + 1e-5, (type(det).__name__, tot_gpu, tot_ref)


def test_runs_above_the_device_counter_are_batched_exactly():
    """McDataTypesSingleCnt64: a budget above ``max_batch`` is run in batches that
    continue the MWC streams and accumulate on the device (one download at the end);
    in deterministic mode the result equals the same batches run one by one with
    ``out=`` - exactly in the integer domain, to the rounding of the per-batch
    conversions in float64 - and the enhanced (two-step) RNG takes part."""
    from pyxopto_b200.mcbase import mctypes
    a = _det_sim('mcml_mhg_gauss_enhanced_rng', types=mctypes.McDataTypesSingleCnt64)[0]
    b = _det_sim('mcml_mhg_gauss_enhanced_rng', types=mctypes.McDataTypesSingleCnt64)[0]
    a.max_batch = 3000
    kw = dict(maxthreads=256, wgsize=64)
    _, flu_a, det_a = a.run(8000, **kw)            # 3000 + 3000 + 2000
    total_a = a.download_raw()[0].copy()
    out = None
    total_b = 0
    for n in (3000, 3000, 2000):
        out = b.run(n, out=out, **kw)
        total_b = total_b + b.download_raw()[0]
    _, flu_b, det_b = out
    assert det_a.top.nphotons == 8000 and flu_a.nphotons == 8000
    assert np.array_equal(total_a, total_b)        # the device holds the exact sum
    assert np.allclose(det_a.top.raw, det_b.top.raw, rtol=1e-14, atol=0) and det_a.top.raw.sum() > 0
    assert np.allclose(flu_a.raw, flu_b.raw, rtol=1e-14, atol=0) and flu_a.raw.sum() > 0
    with pytest.raises(ValueError):
        _det_sim('mcml_mhg_gauss_enhanced_rng')[0].run(2**33)


@pytest.mark.parametrize('name', ['mcml_c1_slab', 'mcvox_gauss_fluence', 'mccyl_hg_line_fiz',
                                  'mcvox_ubeam_radial'])
def test_packet_counter_does_not_wrap_at_the_top_of_its_range(name):
    """A throughput-mode launch of ``max_batch`` packets whose device counter
    starts 20 000 below the budget: the claims of the last chunks (and the one
    extra claim of every thread after exhaustion) must neither wrap the 32-bit
    counter nor drop the tail chunk - the kernel ends, the counter stays above
    the budget and exactly 20 000 packets are simulated (all-bin total within
    5 sigma of a plain 20 000-packet run)."""
    n = 20000
    ref_sim, _, _ = build_sim(name)
    top, _, _ = build_sim(name)
    ref_sim.run(n, download=False)
    a_ref = ref_sim.download_raw()[0].astype(np.float64)
    top._packet_counter_start = top.max_batch - n
    top.run(top.max_batch, download=False)
    cnt = np.zeros(4, np.uint32)
    top._cl_buffers[top._counters_name()].download(top._stream, cnt)
    assert top.max_batch <= int(cnt[0]) <= 0xFFFFFFFF
    assert int(cnt[0]) - top.max_batch <= top.run_report['launched_threads']*top.chunk_max
    a_top = top.download_raw()[0].astype(np.float64)
    K = 0x7FFFFF
    owners = [d for d in (top.detectors or ())] + ([top.fluence] if top.fluence else [])
    for owner in owners:
        scale = float(owner.k) if owner is top.fluence else K
        for al in top.cl_rw_accumulator_allocator.allocations(owner):
            t_top = a_top[al.offset:al.offset + al.size].sum()/scale/n
            t_ref = a_ref[al.offset:al.offset + al.size].sum()/scale/n
            sigma = np.sqrt(max(t_ref, 1e-6)/n)*2
            assert abs(t_top - t_ref) <= 5*sigma + 1e-5, (type(owner).__name__, t_top, t_ref)


def test_device_side_grid_conversion_is_bit_identical():
    """Large fluence grids are converted to float64 on the device (AccuScale)
    instead of by ``update_data`` on the host: same IEEE operations, same bits;
    accumulation with ``out=`` included."""
    def run(on_device):
        sim = _det_sim('mcvox_gauss_fluence')[0]
        sim.SCALE_ON_DEVICE_MIN = 1 if on_device else 1 << 62
        kw = dict(maxthreads=256, wgsize=64)
        res = sim.run(3000, **kw)
        res = sim.run(2000, out=res, **kw)
        return res[1]
    a, b = run(True), run(False)
    assert a.nphotons == b.nphotons == 5000
    assert a.raw.sum() > 0 and a.raw.dtype == np.float64 and a.raw.shape == b.raw.shape
    assert np.array_equal(a.raw, b.raw)


def test_device_side_sampling_volume_conversion_is_bit_identical():
    """The same for the sampling-volume grid (config 4: 200^3 voxels, 64 MB)."""
    name = 'mcml_lut_iso_radialpl_trace'
    out = []
    for on_device in (True, False):
        sim, geom, mc = _det_sim(name)
        sim.SCALE_ON_DEVICE_MIN = 1 if on_device else 1 << 62
        n = run_size(name)[0]
        trace, _, _ = sim.run(n, maxthreads=256, wgsize=64)
        sv = cases.make_sv(mc, name)
        sim.sampling_volume(trace, sv)
        sim.sampling_volume(trace, sv)             # accumulates into the same object
        out.append(sv)
    a, b = out
    assert a.data.sum() > 0 and a.weight == b.weight
    assert np.array_equal(a.data, b.data)


def test_lazy_trace_rows_equal_eager_rows():
    """Rows accepted by the device-side filter are fetched when ``trace.data`` is
    first read; a result that is still outstanding when the next run reuses the
    device buffer is materialised first.  Same rows as the eager download."""
    name = 'mcml_lut_iso_radialpl_trace'
    n = run_size(name)[0]*4
    out = {}
    for lazy in (True, False):
        sim, geom, mc = _det_sim(name)
        sim.trace.filter = _filters(mc)[name]
        sim.lazy_trace_rows = lazy
        kw = dict(maxthreads=256, wgsize=64)
        t1 = sim.run(n, **kw)[0]
        assert t1.rows_on_device_only == lazy
        t2 = sim.run(n, **kw)[0]              # reuses the compact-row buffer
        assert not t1.rows_on_device_only     # ... after t1 was materialised
        sv = cases.make_sv(mc, name)
        sim.sampling_volume(t2, sv)           # consumes the device-resident rows
        assert t2.rows_on_device_only == lazy
        out[lazy] = (t1.data.copy(), t1.n.copy(), t2.data.copy(), t2.n.copy(), sv.data.copy())
    for a, b in zip(out[True], out[False]):
        assert a.shape == b.shape and a.size > 0
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_large_fluence_results_own_their_pinned_grid():
    """Grids converted on the device land in a pooled page-locked array that a fresh
    result keeps without a copy; a later result must not alias an earlier one that is
    still alive, and a released buffer is reused."""
    sim = _det_sim('mcvox_gauss_fluence')[0]
    sim.SCALE_ON_DEVICE_MIN = 1
    kw = dict(maxthreads=256, wgsize=64)
    _, flu_a, _ = sim.run(2000, **kw)
    keep = np.array(flu_a.raw, copy=True)
    _, flu_b, _ = sim.run(2000, **kw)
    assert not np.shares_memory(flu_a.raw, flu_b.raw)
    assert np.array_equal(flu_a.raw, keep) and flu_a.raw.sum() > 0
    addr_a = flu_a.raw.__array_interface__['data'][0]
    del flu_a
    _, flu_c, _ = sim.run(2000, **kw)
    _, flu_d, _ = sim.run(2000, **kw)
    addrs = {f.raw.__array_interface__['data'][0] for f in (flu_b, flu_c, flu_d)}
    assert len(addrs) == 3 and addr_a in addrs          # the released buffer came back
    _, flu_e, _ = sim.run(2000, **kw)                   # pool exhausted: plain copy
    assert not any(np.shares_memory(flu_e.raw, f.raw) for f in (flu_b, flu_c, flu_d))


@pytest.mark.parametrize('name, method', [
    ('mcvox_gauss_fluence', 'aw'), ('mcvox_isopoint_fluencerate', 'aw'),
    ('mcvox_isovoxel_fluence', 'aw'), ('mcvox_ufiber_fluence', 'aw'),
    ('mcvox_gk2_line_total', 'aw'), ('mcvox_ubeam_radial', 'aw'),
    ('mcvox_gauss_fluence', 'ar'), ('mcvox_gk2_line_total', 'ar'),
    ('mcvox_gauss_fluence', 'aw:rmax'), ('mcvox_ubeam_radial', 'aw:rmax')])
def test_both_mcvox_throughput_loops_against_the_oracle(name, method):
    """The packet-pool loop (mcvox_pool_loop.cuh, the default where it applies) and the
    lane-resident loop (mcvox_dda_loop.cuh, ``pool_slots = 0``) are both pinned against the
    oracle: totals within 4 sigma, every bin of the marginal fluence profiles and every
    detector bin within 5 sigma; the two loops must also agree with each other."""
    n = 300000
    K = 0x7FFFFF
    results = {}
    from pyxopto_b200.mcbase import mcoptions
    method, _, rmax = method.partition(':')
    # 64: the pool with its rings of slot numbers (default), 'census': the pool with a census
    # of the slot states and a gather by rank, 0: the lane-resident loop
    for slots in (64, 'census', 0):
        sim, geom, _ = build_sim(name, options=[getattr(mcoptions.McMethod, method)])
        if rmax:
            # an rmax sphere the packets do reach (termination by the end-of-trip test)
            sim.rmax = 0.12e-3
            assert sim._rmax_needed()
        sim.pool_slots = 64 if slots else 0
        sim.pool_queues = slots != 'census'
        sim.run(n, download=False)
        assert sim.run_report['loop'] == ('packet pool' if slots else 'lane-resident rays')
        assert ('#define XO_POOL_QUEUES 1' in sim._last_src) == (slots != 'census')
        results[slots] = sim.download_raw()[0]
    desc = xo_oracle.describe(sim, geom)
    ref = xo_oracle.run(desc, n, 64, sim.rng_seeds_x[:64], sim.rng_seeds_a[:64],
                        math=xo_oracle.MATH_LIBM, method=method)['accu']

    def check(a, b, scale, what, nsig):
        a = a.astype(np.float64)/scale/n
        b = b.astype(np.float64)/scale/n
        sigma = np.sqrt(np.maximum(b, 1e-6)/n)*np.sqrt(2)
        bad = np.abs(a - b) > nsig*sigma + 2e-5
        assert not bad.any(), (what, np.flatnonzero(bad)[:5], a[bad][:5], b[bad][:5])

    rate = bool(sim.resolved_options().get('MC_FLUENCE_MODE_RATE'))
    for slots, other in ((64, ref), ('census', ref), (0, ref), (64, results[0])):
        accu = results[slots]
        for det in sim.detectors or ():
            for a in sim.cl_rw_accumulator_allocator.allocations(det):
                sl = slice(a.offset, a.offset + a.size)
                check(accu[sl].sum(keepdims=True), other[sl].sum(keepdims=True), K,
                      (slots, type(det).__name__, 'total'), 4)
                check(accu[sl], other[sl], K, (slots, type(det).__name__), 5)
        flu = sim.fluence
        if flu is None:
            continue
        for a in sim.cl_rw_accumulator_allocator.allocations(flu):
            g = accu[a.offset:a.offset + a.size].reshape(a.shape)
            c = other[a.offset:a.offset + a.size].reshape(a.shape)
            k = float(flu.k)
            if rate:
                k *= 1.0/min(m.mua for m in sim.materials if m.mua > 0)
            check(g.sum(keepdims=True).ravel(), c.sum(keepdims=True).ravel(), k,
                  (slots, 'fluence total'), 4)
            for axis in range(g.ndim):
                rest = tuple(i for i in range(g.ndim) if i != axis)
                check(g.sum(axis=rest), c.sum(axis=rest), k, (slots, 'fluence', axis), 5)


@pytest.mark.parametrize('slots', [64, 0])
def test_voxel_trace_rows_are_consistent(slots):
    """Full trace of the voxel geometry in throughput mode - packet pool and lane-resident
    loop: every row is a coherent trajectory (launch event first, unit directions, positions
    inside the voxel box, each event on the ray the previous one left along its recorded
    direction, optical path length growing by n x distance, zero tail)."""
    import benchcfg
    from pyxopto_b200.mcvox import mc
    n, maxlen = 20000, 128
    sim = benchcfg.c4_trace_vox(mc, maxlen=maxlen)
    sim.pool_slots = slots
    sim.pool_full_trace = True
    sim.device_trace_filter = False          # raw rows wanted: zero-filled tails
    sim.run(n, download=False)
    assert sim.run_report['loop'] == ('packet pool' if slots else 'lane-resident rays')
    accu, ints, floats = sim.download_raw()
    tp = sim._packed['trace']
    co, do = int(tp.count_buffer_offset), int(tp.data_buffer_offset)
    cnt = ints[co:co + n]
    rows = floats[do:do + n*maxlen*8].reshape(n, maxlen, 8).astype(np.float64)
    assert cnt.min() >= 2
    nrec = np.minimum(cnt, maxlen)
    idx = np.arange(maxlen)[None, :]
    valid = idx < nrec[:, None]
    assert not rows[~valid].any()
    dn = np.linalg.norm(rows[..., 3:6], axis=2)
    assert np.allclose(dn[valid], 1.0, atol=1e-4)
    assert np.allclose(rows[:, 0, :3], 0.0) and np.allclose(rows[:, 0, 5], 1.0)
    half, depth = 201/2*5e-6, 201*5e-6
    pos = rows[..., :3]
    tol = 1e-9
    assert np.abs(pos[..., 0][valid]).max() <= half + tol and np.abs(pos[..., 1][valid]).max() <= half + tol
    assert pos[..., 2][valid].min() >= -tol and pos[..., 2][valid].max() <= depth + tol
    # consecutive events of packets that did not overflow: event i + 1 lies on the ray that
    # leaves event i along the direction recorded there
    whole = cnt <= maxlen
    pair = valid[:, 1:] & valid[:, :-1] & whole[:, None]
    disp = pos[:, 1:] - pos[:, :-1]
    d0 = rows[:, :-1, 3:6]
    dist = np.linalg.norm(disp, axis=2)
    cross = np.linalg.norm(np.cross(disp, d0), axis=2)
    along = (disp*d0).sum(axis=2)
    assert (cross[pair] <= 2e-4*dist[pair] + 2e-10).all()
    assert (along[pair] >= -2e-10).all()
    # optical path length: n x geometric distance (one refractive index in this medium)
    dpl = rows[:, 1:, 7] - rows[:, :-1, 7]
    assert (dpl[pair] >= -1e-9).all()
    assert np.allclose(dpl[pair], 1.337*dist[pair], rtol=2e-3, atol=2e-8)
    # weights never grow except by the survival lottery (x 10)
    w = rows[..., 6]
    ratio = w[:, 1:][pair]/np.maximum(w[:, :-1][pair], 1e-30)
    assert ((ratio <= 1.0 + 1e-6) | (np.abs(ratio - 10.0) <= 1e-4) | (w[:, 1:][pair] == 0.0)).all()
