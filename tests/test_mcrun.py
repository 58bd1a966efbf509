"""Batch runners (``pyxopto_b200.<geometry>.mcrun``): the callers of ``Mc.run`` that the
reference keeps in ``xopto/mcbase/mcrun.py`` and ``xopto/<geometry>/mcrun/mcrun.py``.

CPU: the loop and its stop criteria with a stand-in simulator.  GPU: a runner equals the
hand-written ``run(out=...)`` chain with the same number of batches, bit for bit."""
import types

import numpy as np
import pytest

from helpers import build_sim


class _Rows(types.SimpleNamespace):
    def __len__(self):
        return self.w.size


class _StandIn:
    """Counts calls; every batch adds ``gain`` to one detector bin and ``rows`` trace rows."""

    def __init__(self, mc, gain=2.0, rows=3, top=True):
        self.calls, self.kwargs = 0, []
        self._mc, self._gain, self._rows, self._top = mc, gain, rows, top

    def run(self, nphotons, out=None, **kw):
        self.calls += 1
        self.kwargs.append((nphotons, kw))
        if out is None:
            top = types.SimpleNamespace(raw=np.zeros(4)) if self._top else \
                self._mc.mcdetector.DetectorDefault()
            dets = types.SimpleNamespace(top=top, bottom=self._mc.mcdetector.DetectorDefault(),
                                         specular=types.SimpleNamespace(raw=np.zeros(1)))
            trace = _Rows(w=np.zeros(0))
            out = (trace, None, dets)
        trace, _, dets = out
        if self._top:
            dets.top.raw[1] += self._gain
        trace.w = np.concatenate([trace.w, np.full(self._rows, 0.25)])
        trace.terminal = {'w': trace.w}
        return out


def test_reference_import_paths_and_names():
    from pyxopto_b200.mcml import mcrun as ml
    from pyxopto_b200.mcvox import mcrun as vox
    from pyxopto_b200.mccyl import mcrun as cyl
    from pyxopto_b200.mcml.mcrun.mcrun import RunMinWeightTop      # noqa: F401
    from pyxopto_b200.mcbase.mcrun import RunMinWeightBase, RunMinPacketsBase
    for mod in (ml, vox):
        for n in ('RunMinWeightTop', 'RunMinWeightBottom', 'RunMinWeightSpecular',
                  'RunMinWeightTrace', 'RunMinPacketsTrace'):
            assert issubclass(getattr(mod, n), (RunMinWeightBase, RunMinPacketsBase)), n
    assert not hasattr(cyl, 'RunMinWeightTop')
    assert cyl.RunMinWeightOuter(1.0, 10).location == 'outer'
    assert ml.RunMinWeightBottom(1.0, 10).location == 'bottom'
    assert ml.RunMinWeightTop.__module__ == 'pyxopto_b200.mcml.mcrun'


def test_weight_criterion_and_min_packets():
    from pyxopto_b200.mcml import mc, mcrun
    r = mcrun.RunMinWeightTop(7.0, 100)
    sim = _StandIn(mc)
    out = r.run(sim, wgsize=64)
    assert sim.calls == 4 and r.n == 400 and r.weight == 8.0
    assert out[2].top.raw[1] == 8.0
    assert sim.kwargs[0] == (100, {'wgsize': 64})
    assert (r.min_weight, r.batch_size, r.selection) == (7.0, 100, slice(None))
    # the launched-packet floor keeps the loop going after the weight is there
    sim = _StandIn(mc)
    r.run(sim, min_packets=650)
    assert sim.calls == 7 and r.n == 700
    # a previous result continues; the loop still simulates until the total is there
    sim2 = _StandIn(mc)
    out = r.run(sim2, out=out)
    assert sim2.calls == 1 and r.weight == 10.0
    # selection: only bin 0 counts - never filled, so a weight of 0 ends at once only for min_weight 0
    r0 = mcrun.RunMinWeightTop(0.0, 10, selection=0)
    sim = _StandIn(mc)
    r0.run(sim)
    assert sim.calls == 1 and r0.weight == 0.0
    with pytest.raises(ValueError):
        mcrun.RunMinWeightTop(1.0, 0)


def test_unused_detector_and_missing_trace_raise():
    from pyxopto_b200.mcml import mc, mcrun
    with pytest.raises(RuntimeError, match='"bottom"'):
        mcrun.RunMinWeightBottom(1.0, 10).run(_StandIn(mc))
    with pytest.raises(RuntimeError, match='"top"'):
        mcrun.RunMinWeightTop(1.0, 10).run(_StandIn(mc, top=False))
    sim = types.SimpleNamespace(run=lambda n, out=None: (None, None, None))
    with pytest.raises(RuntimeError, match='trace'):
        mcrun.RunMinWeightTrace(1.0, 10).run(sim)
    with pytest.raises(RuntimeError, match='detector'):
        mcrun.RunMinWeightSpecular(1.0, 10).run(sim)


def test_trace_criteria():
    from pyxopto_b200.mcml import mc, mcrun
    r = mcrun.RunMinWeightTrace(2.0, 50)               # 0.75 of weight per batch
    sim = _StandIn(mc)
    out = r.run(sim)
    assert sim.calls == 3 and r.weight == 2.25 and len(out[0]) == 9
    r = mcrun.RunMinPacketsTrace(10, 50)               # 3 rows per batch (reference: AttributeError)
    sim = _StandIn(mc)
    out = r.run(sim)
    assert sim.calls == 4 and len(out[0]) == 12 and r.n == 200 and r.min_packets == 10


@pytest.mark.gpu
@pytest.mark.parametrize('name, runner, location', [
    ('mcml_c1_slab', 'RunMinWeightTop', 'top'),
    ('mcvox_gauss_fluence', 'RunMinWeightTop', 'top'),
    ('mccyl_gk_ubeam_fiz_trace', 'RunMinWeightOuter', 'outer')])
def test_runner_equals_the_manual_chain(name, runner, location):
    """Deterministic mode: the runner stops after k batches; k manual ``run(out=...)``
    calls on a fresh simulator give the same detector accumulators, and k - 1 batches
    are below the requested weight."""
    import importlib
    from pyxopto_b200.mcbase import mcoptions
    n = 4000
    kw = dict(maxthreads=1024, wgsize=64)
    sim, geom, mc = build_sim(name, options=[mcoptions.McDeterministic.on])
    mcrun = importlib.import_module('pyxopto_b200.{}.mcrun'.format(geom))
    one = float(np.sum(getattr(sim.run(n, **kw)[2], location).raw))
    assert one > 0
    sim = build_sim(name, options=[mcoptions.McDeterministic.on])[0]
    r = getattr(mcrun, runner)(2.5*one, n)
    out = r.run(sim, **kw)
    k = r.n // n
    assert k >= 2 and r.weight >= 2.5*one
    ref_sim = build_sim(name, options=[mcoptions.McDeterministic.on])[0]
    ref, below = None, None
    for i in range(k):
        ref = ref_sim.run(n, out=ref, **kw)
        if i == k - 2:
            below = float(np.sum(getattr(ref[2], location).raw))
    assert below < 2.5*one
    a, b = getattr(out[2], location).raw, getattr(ref[2], location).raw
    assert a.tobytes() == b.tobytes() and getattr(out[2], location).nphotons == k*n
    if out[1] is not None:
        assert np.asarray(out[1].raw).tobytes() == np.asarray(ref[1].raw).tobytes()


@pytest.mark.gpu
def test_packets_trace_runner_on_the_device():
    """``RunMinPacketsTrace`` with a filtered trace: batches until the filter has let
    enough packets through; the rows are the concatenation of the batches' rows."""
    from pyxopto_b200.mcml import mcrun
    from pyxopto_b200.mcbase import mcoptions
    name = 'mcml_lut_iso_radialpl_trace'
    sim, geom, mc = build_sim(name, options=[mcoptions.McDeterministic.on])
    sim.trace.filter = mc.mctrace.Filter(z=(-1.0, 0.0), pz=(-1.0, 0.0))   # left through the top
    n = 800
    r = mcrun.RunMinPacketsTrace(60, n)
    out = r.run(sim, maxthreads=256, wgsize=64)
    tr = out[0]
    assert len(tr) >= 60 and r.n % n == 0 and r.n >= n
    assert tr.data.shape[0] == tr.n.size == len(tr)
    assert np.all(tr.terminal['z'] <= 0.0)
    assert r.weight == pytest.approx(float(np.sum(tr.terminal['w'])))


# ---------------------------------------------------------------------------
# McRunHelper (xopto/mcml/mcrun/helper.py): run-script scaffold
def _skin_helper(musr_values, **mc_kwargs):
    from pyxopto_b200.mcml import mc
    from pyxopto_b200.mcml.mcrun import McRunHelper

    class Helper(McRunHelper):
        sample = -1

        def create_layers(self):
            air = dict(d=float('inf'), mua=0.0, mus=0.0, n=1.0, pf=mc.mcpf.Hg(0.0))
            return mc.mclayer.Layers([
                mc.mclayer.Layer(**air),
                mc.mclayer.Layer(d=2e-3, mua=1e2, mus=100e2, n=1.33, pf=mc.mcpf.Hg(0.8)),
                mc.mclayer.Layer(**air)])

        def create_detectors(self):
            A = mc.mcdetector.Axis
            return mc.mcdetector.Detectors(top=mc.mcdetector.Radial(A(0.0, 2e-3, 50)),
                                           bottom=mc.mcdetector.Total())

        def create_fluence(self):
            A = mc.mcfluence.Axis
            # (square: FluenceRz.data of the reference broadcasts (nr, nz)*(1, nr),
            #  fluencerz.py:343-347 - mirrored, so nr == nz here)
            return mc.mcfluence.FluenceRz(A(0.0, 1e-3, 30), A(0.0, 2e-3, 30))

        def update_layers(self):
            self.sample += 1
            self.mc_obj.layers[1].mus = musr_values[self.sample]/(1.0 - 0.8)
    return Helper(**mc_kwargs)


def test_helper_cli_and_hooks():
    from pyxopto_b200.mcml.mcrun import McRunHelper
    from pyxopto_b200.mcml.mcrun.helper import McRunHelper as same
    assert same is McRunHelper
    opts = McRunHelper.cli_input(n=7, argv=['-f', '3', '-p', '2e6', '-v', '-d', 'B200'])
    assert opts['first'] == 3 and opts['n'] == 7 and opts['packets'] == 2000000
    assert opts['verbose'] is True and opts['device'] == 'B200' and opts['mc_dir'] == 'mc'
    h = _skin_helper([10e2, 20e2])
    cfg = h.collect_mc_config()
    assert set(cfg) == {'source', 'surface', 'layers', 'detectors', 'fluence', 'trace',
                        'rmax', 'run_report'}
    assert cfg['surface'] is None and cfg['trace'] is None and cfg['fluence']['type'] == 'FluenceRz'
    h.update()
    assert h.sample == 0 and h.mc_obj.layers[1].mus == pytest.approx(10e2/0.2)
    assert h.collect_detectors(None) == {'reflectance': None, 'transmittance': None,
                                         'specular': None}


def test_results_from_an_accumulator_row():
    """One row of raw accumulators becomes the result objects ``Mc.run`` returns: the
    blocks of the pack order, scaled by 1/k in float64."""
    from pyxopto_b200 import mcsweep
    h = _skin_helper([10e2])
    sim = h.mc_obj
    sim._pack(1000)
    size = sim.cl_rw_accumulator_allocator.size
    row = (np.arange(size, dtype=np.uint64)*np.uint64(977)) % np.uint64(10007)
    trace, flu, det = mcsweep.results_from_row(sim, row, 1000)
    assert trace is None and det.top.nphotons == flu.nphotons == 1000
    a_top = sim.cl_rw_accumulator_allocator.allocations(sim.detectors.top)[0]
    a_flu = sim.cl_rw_accumulator_allocator.allocations(sim.fluence)[0]
    k = sim.types.mc_accu_k
    assert np.array_equal(det.top.raw, row[a_top.offset:a_top.offset + 50]*(1.0/k))
    assert flu.raw.shape == sim.fluence.shape
    assert np.array_equal(flu.raw.ravel(), row[a_flu.offset:a_flu.offset + 900]*(1.0/flu.k))
    assert type(det.specular).__name__ == 'DetectorDefault'
    assert det.top.reflectance.shape == (50,)


@pytest.mark.gpu
def test_helper_batch_streams_through_the_sweep_driver():
    """Deterministic mode: ``run_batch`` through the sweep driver equals the reference's
    loop of ``run_one`` calls sample by sample (both continue the MWC states from one
    sample to the next), and the recorded configuration is the sample's."""
    from pyxopto_b200.mcbase import mcoptions
    musr = [5e2, 10e2, 20e2, 40e2, 15e2]
    kw = dict(maxthreads=1024, wgsize=64)
    batches = []
    for streamed in (True, False):
        h = _skin_helper(musr, options=[mcoptions.McDeterministic.on], rnginit=97531)
        h.streamed = streamed
        batches.append(h.run_batch(len(musr), 5000, first=100, **kw))
        assert (h._sweep is not None) == streamed
    for i, (a, b) in enumerate(zip(*batches)):
        assert a['index'] == b['index'] == 100 + i and a['num_packets'] == 5000
        assert a['mc']['layers'] == b['mc']['layers']
        assert a['mc']['layers']['layers'][1]['mus'] == pytest.approx(musr[i]/0.2)
        for key in ('reflectance', 'transmittance'):
            assert a['detectors'][key].sum() > 0
            assert np.array_equal(a['detectors'][key], b['detectors'][key]), (i, key)
        assert a['detectors']['specular'] is None
        assert np.array_equal(a['fluence']['data'], b['fluence']['data'])
    r = [d['detectors']['reflectance'].sum() for d in batches[0]]
    assert r[3] > r[0]                      # more scattering, more diffuse reflectance
