"""User-written plugins (OpenCL-C fragments) - CPU part: the reference kernel
executing the fragments of tests/user_plugins.py (golden vectors) agrees with the
built-in plugins it restates and with the oracle; the host mirror packs the user
structs byte-identically; the fragments compile for sm_100a through
csrc/kernels/xo_clcompat*.cuh in both math modes.  The GPU part is
test_gpu_parity.py::test_user_fragments_*."""
import numpy as np
import pytest

import cases
import xo_oracle
from helpers import build_sim, golden, packed_bytes, run_size


def test_reference_runs_user_fragments_like_builtins():
    """Golden vectors of the reference kernel: user Hg + pencil + ring detector ==
    built-in Hg + Line + Radial, accumulators and advanced MWC states."""
    gu, gn = golden('mcml_user_plugins'), golden('mcml_user_plugins_native')
    assert gu['accu'].sum() > 0
    assert np.array_equal(gu['accu'], gn['accu'])
    assert np.array_equal(gu['rng_x_after'], gn['rng_x_after'])


def test_reference_runs_a_user_surface_layout_like_the_builtin():
    """Golden vectors of the reference kernel: the user-written reflector == LambertianReflector."""
    gu, gn = golden('mcml_user_surface_reflector'), golden('mcml_surface_lambert_top')
    assert gu['accu'].sum() > 0
    assert np.array_equal(gu['accu'], gn['accu'])
    assert np.array_equal(gu['rng_x_after'], gn['rng_x_after'])
    assert gu['packed_surface_layouts'].tobytes() == gn['packed_surface_layouts'].tobytes()


@pytest.mark.parametrize('name', ['mcml_user_trace', 'mcvox_user_trace', 'mccyl_user_trace'])
def test_reference_runs_a_user_trace_like_the_builtin(name):
    """Golden vectors of the reference kernel: the user-written trace == Trace, event rows
    and counts included - in every geometry."""
    gu, gn = golden(name), golden(cases.USER_EQUIVALENT[name])
    for key in ('accu', 'ints', 'floats', 'rng_x_after'):
        assert np.array_equal(gu[key], gn[key]), key
    assert gu['ints'].sum() > 0


@pytest.mark.parametrize('name', ['mcml_user_plugins', 'mcml_user_surface_reflector',
                                  'mcml_user_trace', 'mcvox_user_trace', 'mccyl_user_trace'])
def test_oracle_pins_user_fragments_through_equivalent_builtins(name):
    eq = cases.USER_EQUIVALENT[name]
    sim, geom, _ = build_sim(eq)
    n, t = run_size(name)
    sim._pack(n)
    res = xo_oracle.run(xo_oracle.describe(sim, geom), n, t, sim.rng_seeds_x[:t],
                        sim.rng_seeds_a[:t], math=xo_oracle.MATH_LIBM)
    g = golden(name)
    assert np.array_equal(res['accu'], g['accu'])
    assert np.array_equal(res['rng_x'][:t], g['rng_x_after'])


@pytest.mark.parametrize('name', sorted(cases.USER_CASES))
def test_user_structs_pack_like_the_reference(name):
    sim, _, _ = build_sim(name)
    sim._pack(run_size(name)[0])
    g = golden(name)
    mine = packed_bytes(sim)
    medium = ('materials', 'voxels') if name.startswith('mcvox') else ('layers',)
    for key in medium + ('source', 'detectors') + \
            (('fluence',) if 'packed_fluence' in g.files else ()) + \
            (('surface_layouts',) if 'packed_surface_layouts' in g.files else ()) + \
            (('trace',) if 'packed_trace' in g.files else ()):
        assert mine[key] == g['packed_' + key].tobytes(), key


@pytest.mark.parametrize('name', sorted(cases.USER_CASES))
@pytest.mark.parametrize('deterministic', [True, False])
def test_user_fragments_compile_for_sm100a(name, deterministic):
    from pyxopto_b200.mcbase import mcoptions
    kw = dict(options=[mcoptions.McDeterministic.on]) if deterministic else {}
    sim, _, _ = build_sim(name, **kw)
    cubin, log, _ = sim.compile(1000, block=64)
    assert len(cubin) > 10000
    assert 'error' not in log.lower()
    src = sim._last_src
    assert '#include "xo_clcompat.cuh"' in src
    if name == 'mcml_user_fluence':
        assert 'mcsim_fluence_deposit_at' in src and 'typedef xo::FluUser XoFluence;' in src
        assert 'typedef xo::PfHg XoPf;' in src
        return
    if name.endswith('_user_trace') and not name.startswith('mcml'):
        assert 'mcsim_trace_event' in src and '#define XO_USER_TRACE 1' in src
        assert 'typedef xo::PfUser XoPf;' not in src
        return
    if name.startswith('mccyl'):
        assert '#include "xo_clcompat_mccyl.cuh"' in src and 'mc_layer_r_outer' in src
        assert 'typedef xo::SrcUser XoSource;' in src and 'typedef xo::DetUserOuter XoDetOuter;' in src
        assert 'typedef xo::PfUser XoPf;' in src
        return
    if name.startswith('mcvox'):
        assert '#include "xo_clcompat_mcvox.cuh"' in src and 'mcsim_voxel_material' in src
        assert 'typedef xo::SrcUser XoSource;' in src and 'typedef xo::DetUserTop XoDetTop;' in src
        assert 'typedef xo::PfUser XoPf;' in src and 'typedef xo::DetTotal XoDetBottom;' in src
        return
    if name.startswith('mcml_user_trace'):
        assert 'mcsim_trace_event' in src and '#define XO_USER_TRACE 1' in src
        assert '#define TRACE_ENTRY_LEN 8' in src and 'typedef xo::PfUser XoPf;' not in src
        return
    if name.startswith('mcml_user_surface'):
        assert 'mcsim_top_surface_layout_handler' in src
        assert 'typedef xo::SurfUserTop XoSurfTop;' in src and 'typedef xo::PfHg XoPf;' in src
        assert ('typedef xo::SurfUserBottom XoSurfBottom;' in src) == (name.endswith('window'))
        return
    assert 'mcsim_pf_sample_angles' in src
    # built-in slots stay hand-written CUDA
    assert ('typedef xo::SrcLine XoSource;' in src) == (name == 'mcml_user_cubic')


def test_user_pf_builds_in_every_geometry():
    """The phase-function slot takes fragments in mcvox and mccyl too."""
    import importlib
    import user_plugins as up
    mc = importlib.import_module('pyxopto_b200.mcvox.mc')
    vox = cases._vox_grid(mc)
    sim = mc.Mc(vox, cases._vox_materials(mc, lambda g: up.user_cubic(mc, g)),
                mc.mcsource.GaussianBeam(50e-6), rnginit=1)
    cubin, _, _ = sim.compile(1000, block=64)
    assert len(cubin) > 10000 and 'typedef xo::PfUser XoPf;' in sim._last_src
    mc = importlib.import_module('pyxopto_b200.mccyl.mc')
    sim = mc.Mc(cases._cyl_layers(mc, up.user_cubic(mc, 0.5)),
                mc.mcsource.Line((-10e-3, 0.0, 0.0), (1.0, 0.0, 0.0)), rnginit=1)
    cubin, _, _ = sim.compile(1000, block=64)
    assert len(cubin) > 10000 and 'typedef xo::PfUser XoPf;' in sim._last_src


def test_fragment_api_surface_compiles():
    """One fragment that touches the wider API of xo_clcompat.cuh (vector helpers,
    interface physics, lookup tables, layer accessors, option macros)."""
    import user_plugins as up
    sim, _, mc = build_sim('mcml_user_plugins')
    src_plugin = sim.source
    extra = '''
inline mc_fp_t user_api_probe(McSim *mcsim){
	mc_point3f_t a = {FP_1, FP_0, FP_0}, b = {FP_0, FP_1, FP_0}, c;
	mc_point2f_t p2 = {FP_0p5, FP_0p25};
	mc_matrix3f_t T = {FP_1, FP_0, FP_0, FP_0, FP_1, FP_0, FP_0, FP_0, FP_1};
	mc_fp_lut_t lut = {FP_0, FP_1, 2, 0};
	mc_fp_t v = FP_0, s, co;
	mc_cross_point3f(&a, &b, &c);
	transform_point3f(&T, &c, &c);
	mc_normalize_point3f(&c);
	mc_mad_point3f(&a, &b, FP_2, &c);
	mc_sincos(FP_HALF_PI, &s, &co);
	v += mc_dot_point3f(&a, &c) + mc_length_point2f(&p2) + mc_distance_point3f(&a, &b);
	v += reflectance(mc_layer_n(mcsim_layer(mcsim, 0)), mc_layer_n(mcsim_top_sample_layer(mcsim)),
		FP_COS_0, mc_layer_cc_top(mcsim_layer(mcsim, 1)));
	v += cos_critical(FP_LITERAL(1.4), FP_1) + mc_pow(FP_2, s) + mc_exp(-co) + mc_log(FP_2);
	v += mc_atan2(s, co) + mc_acos(FP_0p5) + mc_tan(FP_0p25) + mc_cbrt(FP_2) + mc_rsqrt(FP_4);
	v += (mc_fp_t)mc_clip(mc_round(v), 0, mcsim_layer_count(mcsim)) + (mc_fp_t)mc_fsign(v);
	if (mcsim_fp_lut_array(mcsim) != 0)
		fp_linear_lut_sample(mcsim_fp_lut_array(mcsim), &lut, FP_0p5, &v);
	refract(&a, &b, FP_1, FP_LITERAL(1.33), &c);
	reflect(&a, &b, &c);
	#if MC_USE_TOP_DETECTOR && !MC_USE_FLUENCE
	v += mcsim_position_r(mcsim) + mcsim_direction_z(mcsim) + mcsim_weight(mcsim);
	#endif
	return v + mcsim_optical_pathlength(mcsim) + (mc_fp_t)mcsim_packet_index(mcsim);
};
'''
    base_impl = type(src_plugin).cl_implementation
    type(src_plugin).cl_implementation = staticmethod(
        lambda mc_: base_impl(mc_).replace(
            'inline void mcsim_launch(McSim *mcsim){',
            extra + 'inline void mcsim_launch(McSim *mcsim){\n'
            '\tif (user_api_probe(mcsim) == FP_LITERAL(123456.0)) mcsim_set_weight(mcsim, FP_0);'))
    cubin, log, _ = sim.compile(1000, block=64)
    assert len(cubin) > 10000 and 'user_api_probe' in sim._last_src


def test_plugin_without_any_implementation_is_rejected():
    import importlib
    mc = importlib.import_module('pyxopto_b200.mcml.mc')

    class Bare(mc.mcpf.Hg):
        cu_type = None
    sim = mc.Mc(cases._layers(mc, Bare(0.5)), mc.mcsource.Line(), rnginit=1)
    with pytest.raises(NotImplementedError):
        sim.compile(100)
