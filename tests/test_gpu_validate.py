"""The reference's own acceptance suite (xopto/mcml/test/validate.py) on the GPU,
throughput mode, against the independent CUDA-MCML result vectors the reference
ships (xopto/mcml/test/reference/*.pkl -> tests/golden/validate_vectors.npz,
tests/golden/make_validate.py), with the reference's own pass criteria."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def vec():
    return dict(np.load(os.path.join(GOLDEN_DIR, 'validate_vectors.npz')))


def fiber_reflectance(r, reflectance, sds, dcore, nsimps=1001):
    """Reflectance collected by fibers of diameter ``dcore`` centred ``sds`` from
    the source, from a radial profile: 2 int acos((r^2 + sds^2 - rc^2)/(2 r sds))
    R(r) r dr over [sds - rc, sds + rc] (composite Simpson on linearly
    interpolated R) - the post-processing of the acceptance tests
    (xopto/util/convolve.py:27-132, host maths outside the accelerated path)."""
    reflectance = np.atleast_2d(np.asarray(reflectance, np.float64))
    rc = 0.5*dcore
    out = np.zeros((reflectance.shape[0], len(sds)))
    for j, d in enumerate(sds):
        x = np.linspace(max(d - rc, 0.0), d + rc, nsimps)
        h = x[1] - x[0]
        wts = np.ones(nsimps)
        wts[1:-1:2], wts[2:-1:2] = 4.0, 2.0
        den = 2.0*d*x
        den[den == 0.0] = np.finfo(np.float64).tiny
        ang = np.arccos(np.clip((x**2 + d**2 - rc**2)/den, -1.0, 1.0))
        for i in range(reflectance.shape[0]):
            # linear interpolation with linear extrapolation at both ends
            f = np.interp(x, r, reflectance[i])
            lo, hi = x < r[0], x > r[-1]
            f[lo] = reflectance[i, 0] + (x[lo] - r[0])*(reflectance[i, 1] - reflectance[i, 0])/(r[1] - r[0])
            f[hi] = reflectance[i, -1] + (x[hi] - r[-1])*(reflectance[i, -1] - reflectance[i, -2])/(r[-1] - r[-2])
            out[i, j] = 2.0*h/3.0*np.sum(wts*ang*f*x)
    return out


def _layers(mc, table):
    L = mc.mclayer.Layer
    return mc.mclayer.Layers([L(d=float(d), n=float(n), mua=float(mua), mus=float(mus),
                                pf=mc.mcpf.Hg(float(g))) for d, n, mua, mus, g in table])


def test_single_layer_line_source_radial_profile(vec):
    """validate.SingleLayerLineSourceRadialProfile (validate.py:219-345): mean
    relative error of the radial reflectance profile < 0.5 % at 1e7 packets."""
    from pyxopto_b200.mcml import mc
    start, stop, n = vec['line_axis']
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(
        mc.mcdetector.RadialAxis(float(start), float(stop), int(n)),
        cosmin=float(vec['line_cosmin'])))
    sim = mc.Mc(_layers(mc, vec['line_layers']), mc.mcsource.Line(), det)
    sim.rmax = float(vec['line_rmax'])
    _, _, res = sim.run(10**7)
    rel = (res.top.reflectance - vec['line_reflectance'])/vec['line_reflectance']*100.0
    print('line source radial profile: mean relative error {:+.3f} %'.format(rel.mean()))
    assert abs(rel.mean()) < 0.5


def _fiber_sweep(vec, key, sweep_layer):
    from pyxopto_b200.mcml import mc
    from pyxopto_b200 import mcsweep
    dcore, dclad, ncore, na = [float(v) for v in vec[key + '_fiber']]
    fib = mc.mcsource.MultimodeFiber(dcore, dclad, ncore, na)
    start, stop, n = vec[key + '_axis']
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(
        mc.mcdetector.RadialAxis(float(start), float(stop), int(n)),
        cosmin=float(vec[key + '_cosmin'])))
    table = vec[key + '_layers'].copy()
    if key == 'double':
        mua1, musr1 = vec['double_top_mua_musr']
        table[1, 2], table[1, 3] = mua1, musr1/(1.0 - table[1, 4])
    sim = mc.Mc(_layers(mc, table), mc.mcsource.UniformFiberNI(fib), det)
    sim.rmax = float(vec[key + '_rmax'])
    g = float(table[sweep_layer, 4])
    cfgs = [{sweep_layer: {'mua': float(mua), 'mus': float(musr/(1.0 - g))}}
            for mua in vec[key + '_mua'] for musr in vec[key + '_musr']]
    nphotons = 10**7
    sweep = mcsweep.Sweep(sim)
    _, rows = sweep.run(cfgs, nphotons)
    top = sim.detectors.top
    raw = sweep.detector(rows, top, nphotons)                      # weight per bin
    refl = raw*top._inv_accumulators_area[None, :]/nphotons
    simulated = fiber_reflectance(top.r, refl, vec[key + '_sds'], float(vec[key + '_dcore']))
    reference = vec[key + '_reflectance']
    rel = (simulated.reshape(reference.shape) - reference)/reference*100.0
    return rel, sweep.report['seconds']


def test_single_layer_uniform_fiber_radial(vec):
    """validate.SingleLayerUniformFiberRadial (validate.py:348-511), the reference's
    performance workload: 400 (mua, musr) points x 1e7 packets; |mean relative
    error| of the fiber reflectances < 0.5 %."""
    rel, seconds = _fiber_sweep(vec, 'single', 1)
    print('single layer uniform fiber: mean relative error {:+.3f} % over {} values; '
          '400 x 1e7 packets in {:.2f} s (reference: 5.7 s on an RTX A6000)'.format(
              rel.mean(), rel.size, seconds))
    assert abs(rel.mean()) < 0.5


def test_double_layer_uniform_fiber_radial(vec):
    """validate.DoubleLayerUniformFiberRadial (validate.py:692-864): per-fiber mean
    relative error < 0.5 % for every source-detector separation."""
    rel, _ = _fiber_sweep(vec, 'double', 2)
    per_fiber = rel.mean(axis=(0, 1))
    print('double layer uniform fiber: mean relative error per fiber', per_fiber)
    assert np.abs(per_fiber).max() < 0.5


def test_single_layer_uniform_fiber_trace(vec):
    """validate.SingleLayerUniformFiberTrace (validate.py:1059-1224): the radial
    reflectance reconstructed from the terminal trace events equals the Radial
    detector's within 0.1 % in every bin."""
    from pyxopto_b200.mcml import mc
    dcore, dclad, ncore, na = [float(v) for v in vec['trace_fiber']]
    fib = mc.mcsource.MultimodeFiber(dcore, dclad, ncore, na)
    start, stop, n = vec['trace_axis']
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(
        mc.mcdetector.RadialAxis(float(start), float(stop), int(n)),
        cosmin=float(vec['trace_cosmin'])))
    trace = mc.mctrace.Trace(maxlen=int(vec['trace_maxlen']))
    sim = mc.Mc(_layers(mc, vec['trace_layers']), mc.mcsource.UniformFiber(fib), det, trace)
    sim.rmax = float(vec['trace_rmax'])
    g = float(vec['trace_layers'][1, 4])
    nphotons = 200000
    nr = sim.detectors.top.n
    r = sim.detectors.top.edges
    dr = r[1] - r[0]
    cosmin = sim.detectors.top.cosmin
    arings = np.pi*(r[1:]**2 - r[:-1]**2)
    pkt = np.arange(nphotons)
    worst = 0.0
    for mua, musr in zip(vec['trace_mua'], vec['trace_musr']):
        sim.layers[1].mua = float(mua)
        sim.layers[1].mus = float(musr/(1.0 - g))
        tr, _, res = sim.run(nphotons)
        d = tr.data
        term = np.minimum(tr.n - 1, tr.maxlen - 1)
        x, y, z = d['x'][pkt, term], d['y'][pkt, term], d['z'][pkt, term]
        pz, w = d['pz'][pkt, term], d['w'][pkt, term]
        rind = np.minimum(np.floor(np.sqrt(x**2 + y**2)/dr), nr - 1)
        mask = np.logical_and(z <= 0.0, np.abs(pz) >= cosmin)
        from_trace = np.bincount(rind[mask].astype(np.int64), weights=w[mask].astype(np.float64),
                                 minlength=nr)/(nphotons*arings)
        radial = res.top.reflectance
        nz = radial != 0
        assert not np.any(from_trace[~nz] != 0.0)
        if nz.any():
            worst = max(worst, float(np.abs((from_trace[nz] - radial[nz])/radial[nz]).max()))
    print('trace vs Radial detector: largest relative difference {:.2e}'.format(worst))
    assert worst*100.0 < 0.1
