"""Binary64 kernels (``McDataTypesDouble``, xopto/mcbase/mctypes.py:647-748,991-1044).

Golden vectors: the REFERENCE kernel rendered in double precision and compiled for
the CPU (oracle/refkernel.py with -DXO_REF_DOUBLE; generated in the container that
holds the reference).  The CUDA kernels are compiled with XO_DOUBLE: the same text in
binary64, reference expression and draw order, static block schedule.

CPU: packed structs byte-identical to the reference's double-precision packing, the
translation units compile for sm_100a.  GPU: accumulators, trace rows and advanced MWC
states against the golden vectors - IEEE operations are identical, CUDA's and glibc's
``log`` / ``sincos`` / ``cbrt`` may differ in the last place of a double, which moves a
fixed-point deposit by one unit at most once in ~1e9 deposits and a branch never in
practice: integer buffers are expected (and required) to be equal, trace rows to 1e-9
relative (north star: 1e-5).
"""
import importlib

import numpy as np
import pytest

import cases
from helpers import golden, packed_bytes


def _sim(name, **kw):
    geom = cases.DOUBLE_GEOMETRY[name]
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geom))
    sim, attrs = cases.DOUBLE_CASES[name](mc, **kw)
    for k, v in attrs.items():
        setattr(sim, k, v)
    return sim, geom, mc


@pytest.mark.parametrize('name', sorted(cases.DOUBLE_CASES))
def test_double_structs_pack_like_the_reference(name):
    sim, _, _ = _sim(name)
    g = golden(name)
    sim._pack(int(g['nphotons']))
    mine = packed_bytes(sim)
    keys = [k for k in g.files if k.startswith('packed_')]
    assert keys
    for key in keys:
        assert mine[key[len('packed_'):]] == g[key].tobytes(), key
    if len(sim._float_lut):
        lut = sim._float_lut.pack_into(None)
        assert lut.dtype == np.float64 and g['lut'].dtype == np.float64
        assert np.array_equal(lut[:g['lut'].size], g['lut'][:lut.size])
    assert np.array_equal(sim.rng_seeds_x[:int(g['nthreads'])], g['rng_x0'])


@pytest.mark.parametrize('name', sorted(cases.DOUBLE_CASES))
def test_double_kernels_compile_for_sm100a(name):
    sim, _, _ = _sim(name)
    cubin, log, _ = sim.compile(1000, block=64)
    assert len(cubin) > 10000
    src = sim._last_src
    assert '#define XO_DOUBLE 1' in src and '#define XO_DETERMINISTIC 1' in src
    assert sim.deterministic        # binary64 runs use the reference-structured loops


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(cases.DOUBLE_CASES))
def test_double_precision_against_the_reference_kernel(name):
    sim, geom, _ = _sim(name)
    g = golden(name)
    n, t = int(g['nphotons']), int(g['nthreads'])
    sim.run(n, maxthreads=t, wgsize=t, download=False)
    assert sim.run_report['launched_threads'] == t
    accu, ints, floats = sim.download_raw()
    assert floats.dtype == np.float64
    assert accu.sum() > 0
    assert np.array_equal(accu, g['accu'])
    assert np.array_equal(ints, g['ints'])
    assert np.array_equal(sim.download_seeds()[:t], g['rng_x_after'])
    # (positions are metres around 1e-3: last-place differences of the elementary
    # functions reach ~1e-16 m after a few hundred events)
    assert np.allclose(floats, g['floats'], rtol=1e-9, atol=1e-15)
    assert sim.run_report['threads'] == int(g['num_kernels'])


@pytest.mark.gpu
def test_double_results_come_back_in_reference_units():
    sim, _, _ = _sim('mcml_double_mhg_gauss_cart_flurz')
    trace, fluence, detectors = sim.run(20000)
    assert fluence.raw.dtype == np.float64 and detectors.top.raw.sum() > 0
    single = importlib.import_module('pyxopto_b200.mcml.mc')
    ref, attrs = cases.mcml_mhg_gauss_cart_flurz(single)
    for k, v in attrs.items():
        setattr(ref, k, v)
    _, flu32, det32 = ref.run(20000)
    # the two precisions are the same physics (different draws): totals within 5 sigma
    a, b = detectors.top.raw.sum()/20000, det32.top.raw.sum()/20000
    assert abs(a - b) < 5*np.sqrt(2*b/20000)
    a, b = fluence.raw.sum()/20000, flu32.raw.sum()/20000
    assert abs(a - b) < 5*np.sqrt(2*b/20000)


def test_sampling_volume_kernel_compiles_in_double():
    from pyxopto_b200.mcbase.mcsim import compile_kernel
    sim, _, _ = _sim('mcml_double_lut_iso_radialpl_trace')
    src = sim._SV_SRC.format(det=1, dbl=1)
    cubin, log, _ = compile_kernel(src, True, arch='sm_100a',
                                   extra_options=sim._cl_build_options)
    assert len(cubin) > 10000 and 'error' not in log.lower()


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(cases.SV_DOUBLE_CASES))
def test_double_sampling_volume_against_the_reference_kernel(name):
    """``Mc.sampling_volume`` in binary64 against the reference's SamplingVolume kernel
    rendered in double precision, fed with the same (reference-kernel) trace rows: the
    64-bit voxel accumulators and the total weight must be equal."""
    sim, geom, mc = _sim(name)
    g = golden(name)
    n, t = int(g['nphotons']), int(g['nthreads'])
    trace, _, _ = sim.run(n, maxthreads=t, wgsize=t)
    assert trace.nphotons == n and trace.data.dtype.itemsize % 8 == 0
    # the rows of the reference kernel (ours agree to 1e-9, tested above)
    P = sim._packed['trace']
    do, co = int(P.data_buffer_offset), int(P.count_buffer_offset)
    ml = int(sim.trace.maxlen)
    assert np.array_equal(np.asarray(trace.n), g['ints'][co:co + n])
    rows = g['floats'][do:do + n*ml*8]
    trace.data.view(np.float64).reshape(-1)[:] = rows
    trace._device_token = None      # (upload these rows instead of reading the resident ones)
    sv = cases.make_sv(mc, cases.SV_DOUBLE_CASES[name])
    sim.sampling_volume(trace, sv, maxthreads=256)
    assert bytes(memoryview(sim._packed['sv']).cast('B')) == g['sv_packed'].tobytes()
    assert bytes(memoryview(sim._packed['sv_trace']).cast('B')) == g['sv_packed_trace'].tobytes()
    accu, _, _ = sim.download_raw()
    assert g['sv_accu'].sum() > 0
    assert np.array_equal(accu[:g['sv_accu'].size], g['sv_accu'])
    assert sv.weight == int(g['sv_total_weight'])/sv.k
    assert sv.data.shape == sv.shape
