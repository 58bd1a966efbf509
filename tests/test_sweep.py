"""Config-5 sweeps (pyxopto_b200/mcsweep.py): static pseudo-random partition (CPU)
and, on the GPU, the pipelined sweep against one blocking ``Mc.run`` per
configuration."""
import numpy as np
import pytest

from pyxopto_b200 import mcsweep


def test_partition_is_disjoint_and_covering():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            seen = np.concatenate([mcsweep.partition(n, world, r) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(n))
            sizes = [mcsweep.partition(n, world, r).size for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_partition_is_disjoint_covering_and_balanced():
    rng = np.random.default_rng(3)
    for n in (1, 9, 4096):
        costs = rng.lognormal(0.0, 0.7, n)
        for world in (1, 2, 8):
            parts = [mcsweep.balanced_partition(costs, world, r) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            sizes = [p.size for p in parts]
            assert max(sizes) - min(sizes) <= 1
            if n == 4096 and world == 8:
                loads = np.array([costs[p].sum() for p in parts])
                assert loads.max()/loads.mean() - 1.0 < 1e-3
                perm = np.array([costs[mcsweep.partition(n, world, r)].sum()
                                 for r in range(world)])
                assert loads.max() < perm.max()


def _configs():
    g = 0.8
    return [{1: {'mua': float(mua), 'mus': float(musr/(1.0 - g))}}
            for mua in np.linspace(0.0, 5e2, 3) for musr in np.linspace(5e2, 35e2, 3)]


def _sweep_sim():
    import benchcfg
    from pyxopto_b200.mcbase import mcoptions
    from pyxopto_b200.mcml import mc
    return benchcfg.c5_slab(mc, options=[mcoptions.McDeterministic.on]), mc


@pytest.mark.gpu
def test_pipelined_sweep_equals_one_run_per_configuration():
    n = 20000
    configs = _configs()
    sim, mc = _sweep_sim()
    sweep = mcsweep.Sweep(sim)
    idx, rows = sweep.run(configs, n, maxthreads=1024, wgsize=64)
    assert idx.tolist() == list(range(len(configs)))
    ref_sim, _ = _sweep_sim()
    for i, cfg in enumerate(configs):
        mcsweep._apply_layer_updates(ref_sim, cfg)
        ref_sim.run(n, maxthreads=1024, wgsize=64, download=False)
        accu = ref_sim.download_raw()[0]
        assert np.array_equal(rows[i], accu), i
    refl = sweep.detector(rows, sim.detectors.top, n)
    assert refl.shape == (len(configs), 500)
    total = refl.sum(axis=1)/n
    assert (total > 0.01).all() and (total < 1.0).all()
    # more absorption -> less diffuse reflectance at fixed scattering
    assert total[0] > total[3] > total[6]
    # rank 1 of 2 simulates its share of the fixed permutation
    sweep2 = mcsweep.Sweep(_sweep_sim()[0], rank=1, world=2)
    idx2, rows2 = sweep2.run(configs, n, maxthreads=1024, wgsize=64)
    assert idx2.tolist() == mcsweep.partition(len(configs), 2, 1).tolist()
    assert 0 < len(idx2) < len(configs)
    assert rows2.shape == (len(idx2), rows.shape[1])


@pytest.mark.gpu
def test_pilot_costs_and_balanced_deal():
    """The pilot run reports loop trips per configuration (more scattering = more
    trips); dealing by those costs changes which rank simulates what, not the rows."""
    n = 20000
    configs = _configs()
    sweep = mcsweep.Sweep(_sweep_sim()[0])
    costs = sweep.pilot_costs(configs, 1000)
    assert costs.shape == (len(configs),) and (costs > 0).all()
    assert costs.max() > 1.2*costs.min()        # the grid points differ in cost
    # (fresh simulators: consecutive runs of one simulator continue its MWC streams)
    idx_a, rows_a = mcsweep.Sweep(_sweep_sim()[0]).run(configs, n, maxthreads=1024, wgsize=64)
    idx_b, rows_b = mcsweep.Sweep(_sweep_sim()[0]).run(configs, n, maxthreads=1024, wgsize=64,
                                                       costs=costs)
    assert idx_a.tolist() == idx_b.tolist()
    assert np.array_equal(rows_a, rows_b)
    # two ranks: the shares are disjoint, covering and follow the costs
    parts = [mcsweep.Sweep(_sweep_sim()[0], rank=r, world=2) for r in range(2)]
    got = [p.run(configs, n, maxthreads=1024, wgsize=64, costs=costs)[0] for p in parts]
    assert sorted(np.concatenate(got).tolist()) == list(range(len(configs)))
    assert got[0].tolist() == mcsweep.balanced_partition(costs, 2, 0).tolist()


@pytest.mark.gpu
def test_two_lane_sweep_in_throughput_mode():
    """Throughput mode runs consecutive configurations on two lanes (own streams,
    accumulators, counters, MWC states and packed tables): every row must be the result of
    ITS configuration - statistically equal to a stand-alone run and to the single-lane
    sweep - and the two lanes must draw from different seed sets."""
    import benchcfg
    from pyxopto_b200.mcml import mc
    n = 400000
    configs = _configs()

    def fast_sim():
        return benchcfg.c5_slab(mc)

    sim = fast_sim()
    sweep = mcsweep.Sweep(sim)
    assert sweep.overlap
    idx, rows = sweep.run(configs, n)
    assert sim._lane == 0 and len(sim._lane_streams) == 2
    assert not np.array_equal(sim._lane_seeds[1][0][:64], sim.rng_seeds_x[:64])
    single = mcsweep.Sweep(fast_sim())
    single.overlap = False
    _, rows_1 = single.run(configs, n)
    assert len(single.sim._lane_streams) == 1
    K = 0x7FFFFF
    alone = fast_sim()
    for i, cfg in enumerate(configs):
        mcsweep._apply_layer_updates(alone, cfg)
        alone.run(n, download=False)
        ref = alone.download_raw()[0]
        for got in (rows[i], rows_1[i]):
            a, b = got.sum()/K/n, ref.sum()/K/n
            sigma = np.sqrt(2*max(b, 1e-6)/n)
            assert abs(a - b) <= 4.5*sigma, (i, a, b)
            assert int(got.sum()) != int(ref.sum())      # (not the same packets)
    assert sweep.report['iterations'].shape == (len(configs),)
    assert (sweep.report['iterations'] > 0).all()
