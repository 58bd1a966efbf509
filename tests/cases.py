"""Named simulator configurations shared by the golden-vector generator (which
builds them with the *reference* package) and the tests (which build them with
``pyxopto_b200``).  The two packages expose the same plugin API, so every case is
a function of the simulator module ``mc`` (``xopto.mcml.mc`` or
``pyxopto_b200.mcml.mc``) - the tests therefore read like reference scripts.
"""
import numpy as np

RNGINIT = 123456789


def _layers(mc, pf, stack='two'):
    L = mc.mclayer.Layer
    if stack == 'slab':      # BASELINE config 1: 10 mm slab, n=1.33, mua 1/cm, mus 100/cm
        return mc.mclayer.Layers([
            L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
            L(d=1e-2, n=1.33, mua=1e2, mus=100e2, pf=pf),
            L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])
    return mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=1e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=2e-3, n=1.4, mua=0.5e2, mus=50e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])


def _fiber(mc):
    try:
        from xopto.mcml.mcutil import fiber as fiberutil   # reference
        if mc.__name__.startswith('xopto'):
            return fiberutil.MultimodeFiber(200e-6, 220e-6, 1.462, 0.22)
    except ImportError:
        pass
    return mc.mcsource.MultimodeFiber(200e-6, 220e-6, 1.462, 0.22)


def _hg_lut():
    """(params, lut) of a 2000-point Hg(0.8) lookup table; generated once with
    the reference's xopto.pf.Hg(0.8).mclut(2000) and stored as a fixture."""
    import os
    f = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lut_hg08_2000.npz'))
    return f['params'], f['lut']


def mcml_c1_slab(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(Axis(0.0, 10e-3, 1000)),
        bottom=mc.mcdetector.Radial(Axis(0.0, 10e-3, 1000)),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8), 'slab'), mc.mcsource.Line(), det,
                 rnginit=RNGINIT, **kw), dict(rmax=float('inf'))


def mcml_hg_line_radial(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(Axis(0, 10e-3, 1000)),
        bottom=mc.mcdetector.Radial(Axis(0, 10e-3, 100), cosmin=0.5),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det,
                 rnginit=RNGINIT, **kw), dict(rmax=20e-3)


def mcml_mhg_gauss_cart_flurz(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Cartesian(Axis(-2e-3, 2e-3, 40)),
        specular=mc.mcdetector.Radial(Axis(0, 1e-3, 10)))
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 40), Axis(0, 3e-3, 60))
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.8, 0.9)), mc.mcsource.GaussianBeam(100e-6), det,
                 fluence=flu, rnginit=987654321, **kw), dict(rmax=20e-3)


def mcml_gk_fiber_six_flu(mc, **kw):
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.SixAroundOne(fib), bottom=mc.mcdetector.Total(),
        specular=mc.mcdetector.Total())
    flu = mc.mcfluence.Fluence(Axis(-1e-3, 1e-3, 20), Axis(-1e-3, 1e-3, 20),
                               Axis(0, 3e-3, 30), mode='deposition')
    return mc.Mc(_layers(mc, mc.mcpf.Gk(0.8, 0.5)), mc.mcsource.UniformFiber(fib), det,
                 fluence=flu, rnginit=5555, **kw), dict(rmax=20e-3)


def mcml_lut_iso_radialpl_trace(mc, **kw):
    Axis = mc.mcdetector.Axis
    params, lut = _hg_lut()
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.RadialPl(Axis(0, 5e-3, 50), Axis(0, 0.05, 100)),
        bottom=mc.mcdetector.TotalPl(Axis(0, 0.05, 100)))
    tr = mc.mctrace.Trace(maxlen=50, options=mc.mctrace.Trace.TRACE_ALL)
    return mc.Mc(_layers(mc, mc.mcpf.Lut(params, lut)),
                 mc.mcsource.IsotropicPoint((0, 0, 0.5e-3)), det, trace=tr,
                 rnginit=777, **kw), dict(rmax=20e-3)


def mcml_hg_line_total_fluencet(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(cosmin=0.9),
                                  bottom=mc.mcdetector.Total())
    flu = mc.mcfluence.Fluencet(Axis(-1e-3, 1e-3, 10), Axis(-1e-3, 1e-3, 10),
                                Axis(0, 3e-3, 15), Axis(0, 50e-12, 20))
    tr = mc.mctrace.Trace(maxlen=2, options=mc.mctrace.Trace.TRACE_START |
                          mc.mctrace.Trace.TRACE_END, plon=True)
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.0)), mc.mcsource.Line((0, 0, 0), (0.3, 0, 1)), det,
                 fluence=flu, trace=tr, rnginit=4242, **kw), dict(rmax=5e-3)


def mcml_surface_six_lambert(mc, **kw):
    """Surface layouts (mcml/mcsurface): a six-around-one probe with a filled
    cut-out and a reflective tip on the top surface, a partly specular
    Lambertian reflector under a thin two-layer sample."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    L = mc.mclayer.Layer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=0.3e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=0.4e-3, n=1.4, mua=0.5e2, mus=50e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])
    surf = mc.mcsurface.SurfaceLayouts(
        top=mc.mcsurface.SixAroundOne(fib, spacing=240e-6, diameter=3e-3, reflectivity=0.6,
                                      cutout=0.9e-3, cutoutn=1.6, position=(20e-6, -10e-6),
                                      direction=(0.05, 0.0, 1.0)),
        bottom=mc.mcsurface.LambertianReflector(reflectance=0.9, specular=0.3))
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.SixAroundOne(fib, spacing=240e-6, position=(20e-6, -10e-6),
                                       direction=(0.05, 0.0, 1.0)),
        bottom=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 40), Axis(0, 0.7e-3, 35))
    return mc.Mc(layers, mc.mcsource.UniformFiber(fib), det, fluence=flu, surface=surf,
                 rnginit=24680, **kw), dict(rmax=20e-3)


def mcml_surface_lambert_top(mc, **kw):
    """Ideal Lambertian reflector on the top surface, open bottom."""
    Axis = mc.mcdetector.Axis
    surf = mc.mcsurface.SurfaceLayouts(top=mc.mcsurface.LambertianReflector(0.8, 0.0))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(),
                                  bottom=mc.mcdetector.Radial(Axis(0, 5e-3, 50)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det, surface=surf,
                 rnginit=13579, **kw), dict(rmax=20e-3)


def mcml_mhg_fiber_cartpl_sixpl(mc, **kw):
    """Path-length resolved (TOF) detectors: CartesianPl on the bottom surface
    (linear pl axis) and SixAroundOnePl on the top surface (logarithmic pl axis)."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.SixAroundOnePl(fib, spacing=230e-6,
                                         plaxis=Axis(1e-4, 0.1, 12, logscale=True)),
        bottom=mc.mcdetector.CartesianPl(Axis(-1.5e-3, 1.5e-3, 12), Axis(-1e-3, 1e-3, 8),
                                         plaxis=Axis(0.0, 0.03, 15), cosmin=0.2),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.7, 0.95)), mc.mcsource.UniformFiber(fib), det,
                 rnginit=1122334455, **kw), dict(rmax=20e-3)


def mcml_hg_ubeam_radial(mc, **kw):
    """Tilted elliptical top-hat beam (mcml UniformBeam), Radial detectors and a
    specular Cartesian detector."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(Axis(0, 4e-3, 80), position=(0.1e-3, 0.0)),
        bottom=mc.mcdetector.Radial(Axis(0, 4e-3, 40)),
        specular=mc.mcdetector.Cartesian(Axis(-1e-3, 1e-3, 10)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.85)),
                 mc.mcsource.UniformBeam((400e-6, 250e-6), position=(0.1e-3, 0.0, 0.0),
                                         direction=(0.2, -0.1, 1.0)),
                 det, rnginit=998877, **kw), dict(rmax=20e-3)


def _pf_case(make_pf, rnginit):
    def case(mc, **kw):
        Axis = mc.mcdetector.Axis
        det = mc.mcdetector.Detectors(
            top=mc.mcdetector.Radial(Axis(0, 5e-3, 100)), bottom=mc.mcdetector.Total(),
            specular=mc.mcdetector.Total())
        return mc.Mc(_layers(mc, make_pf(mc)), mc.mcsource.Line((0, 0, 0), (0.1, 0.05, 1.0)),
                     det, rnginit=rnginit, **kw), dict(rmax=20e-3)
    return case


# the remaining built-in phase functions (mcbase/mcpf: hg2, gk2, mgk, pc, mpc)
mcml_pf_hg2 = _pf_case(lambda mc: mc.mcpf.Hg2(0.9, -0.3, 0.15), 31)
mcml_pf_gk2 = _pf_case(lambda mc: mc.mcpf.Gk2(0.85, 0.6, -0.4, 0.5, 0.2), 32)
mcml_pf_mgk = _pf_case(lambda mc: mc.mcpf.MGk(0.8, 0.7, 0.9), 33)
mcml_pf_pc = _pf_case(lambda mc: mc.mcpf.Pc(6.0), 34)
mcml_pf_mpc = _pf_case(lambda mc: mc.mcpf.MPc(8.0, 0.85), 35)


def mcml_hg_gauss_fluencerzt(mc, **kw):
    """Time-resolved r-z deposition (FluenceRzt)."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRzt(Axis(0, 2e-3, 20), Axis(0, 3e-3, 30), Axis(0, 40e-12, 16),
                                  center=(0.1e-3, 0.0))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.GaussianBeam(150e-6), det,
                 fluence=flu, rnginit=60606, **kw), dict(rmax=20e-3)


def mcml_hg_fiber_fluencecyl(mc, **kw):
    """Cylindrical r-fi-z deposition grid (FluenceCyl)."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceCyl(Axis(0, 2e-3, 20), Axis(0, 2*np.pi, 12), Axis(0, 3e-3, 30),
                                  center=(0.05e-3, -0.05e-3))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.UniformFiber(fib, position=(0.1e-3, 0, 0), direction=(0.2, 0.1, 1)),
                 det, fluence=flu, rnginit=70707, **kw), dict(rmax=20e-3)


def mcml_hg_line_symmetricx(mc, **kw):
    """SymmetricX detectors: linear bins on top, logarithmic bins at the bottom."""
    SA = mc.mcdetector.SymmetricAxis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.SymmetricX(SA(0.1e-3, 3e-3, 30), cosmin=0.3),
        bottom=mc.mcdetector.SymmetricX(SA(0.0, 3e-3, 20, logscale=True)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line((0.05e-3, 0, 0), (0.1, 0, 1)),
                 det, rnginit=80808, **kw), dict(rmax=20e-3)


MCML_CASES = {
    'mcml_hg_line_symmetricx': mcml_hg_line_symmetricx,
    'mcml_hg_gauss_fluencerzt': mcml_hg_gauss_fluencerzt,
    'mcml_hg_fiber_fluencecyl': mcml_hg_fiber_fluencecyl,
    'mcml_pf_hg2': mcml_pf_hg2, 'mcml_pf_gk2': mcml_pf_gk2, 'mcml_pf_mgk': mcml_pf_mgk,
    'mcml_pf_mpc': mcml_pf_mpc,
    'mcml_hg_ubeam_radial': mcml_hg_ubeam_radial,
    'mcml_mhg_fiber_cartpl_sixpl': mcml_mhg_fiber_cartpl_sixpl,
    'mcml_surface_six_lambert': mcml_surface_six_lambert,
    'mcml_surface_lambert_top': mcml_surface_lambert_top,
    'mcml_c1_slab': mcml_c1_slab,
    'mcml_hg_line_radial': mcml_hg_line_radial,
    'mcml_mhg_gauss_cart_flurz': mcml_mhg_gauss_cart_flurz,
    'mcml_gk_fiber_six_flu': mcml_gk_fiber_six_flu,
    'mcml_lut_iso_radialpl_trace': mcml_lut_iso_radialpl_trace,
    'mcml_hg_line_total_fluencet': mcml_hg_line_total_fluencet,
}

ALL_CASES = dict(MCML_CASES)
GEOMETRY = {name: name.split('_')[0] for name in ALL_CASES}

# (packets, work-items) of the static block schedule used for the golden vectors
GOLDEN_RUN = {name: (3000, 16) for name in ALL_CASES}
GOLDEN_RUN['mcml_c1_slab'] = (4000, 64)
GOLDEN_RUN['mcml_lut_iso_radialpl_trace'] = (800, 16)


# ---------------------------------------------------------------------------
# voxelised cases
def _vox_grid(mc, n=(24, 20, 28), voxel=25e-6):
    A = mc.mcgeometry.Axis
    nx, ny, nz = n
    return mc.mcgeometry.Voxels(A(-nx/2*voxel, nx/2*voxel, nx), A(-ny/2*voxel, ny/2*voxel, ny),
                                A(0.0, nz*voxel, nz))


def _vox_materials(mc, pf_factory, n_vessel=1.337):
    M = mc.mcmaterial.Material
    return mc.mcmaterial.Materials([
        M(n=1.0, mua=0.0, mus=0.0, pf=pf_factory(1.0)),
        M(n=1.337, mua=16.5724e2, mus=375.9398e2, pf=pf_factory(0.9)),
        M(n=1.4, mua=0.4585e2, mus=356.5406e2, pf=pf_factory(0.9)),
        M(n=n_vessel, mua=230.5427e2, mus=93.9850e2, pf=pf_factory(0.9))])


def _fill_skin_vessel(sim, depth=100e-6, center=350e-6, radius=120e-6):
    z, y, x = sim.voxels.meshgrid()
    m = sim.voxels.material
    m[z <= depth] = 1
    m[z > depth] = 2
    m[(x**2 + (z - center)**2) <= radius**2] = 3
    return sim


def mcvox_gauss_fluence(mc, **kw):
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc)
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(A(0, 0.4e-3, 40)), bottom=mc.mcdetector.Total(),
        specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg), mc.mcsource.GaussianBeam(50e-6),
                detectors=det, fluence=flu, rnginit=2468, **kw)
    return _fill_skin_vessel(sim), dict(rmax=25e-3)


def mcvox_line_mhg_trace(mc, **kw):
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc, n=(16, 16, 20))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Cartesian(A(-0.2e-3, 0.2e-3, 16)),
                                  bottom=mc.mcdetector.Radial(A(0, 0.3e-3, 10), cosmin=0.3))
    tr = mc.mctrace.Trace(maxlen=40, options=mc.mctrace.Trace.TRACE_ALL, plon=True)
    sim = mc.Mc(vox, _vox_materials(mc, lambda g: mc.mcpf.MHg(g, 0.85), n_vessel=1.36),
                mc.mcsource.Line((10e-6, -5e-6, 0.0), (0.2, 0.1, 1.0)),
                detectors=det, trace=tr, rnginit=1357, **kw)
    return _fill_skin_vessel(sim, center=250e-6, radius=80e-6), dict(rmax=5e-3)


def mcvox_line_mhg_trace_startend(mc, **kw):
    """Launch and terminal events only (TRACE_START | TRACE_END) in the voxel geometry."""
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc, n=(16, 16, 20))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Cartesian(A(-0.2e-3, 0.2e-3, 16)),
                                  bottom=mc.mcdetector.Radial(A(0, 0.3e-3, 10), cosmin=0.3))
    T = mc.mctrace.Trace
    tr = T(maxlen=2, options=T.TRACE_START | T.TRACE_END, plon=True)
    sim = mc.Mc(vox, _vox_materials(mc, lambda g: mc.mcpf.MHg(g, 0.85), n_vessel=1.36),
                mc.mcsource.Line((10e-6, -5e-6, 0.0), (0.2, 0.1, 1.0)),
                detectors=det, trace=tr, rnginit=2468, **kw)
    return _fill_skin_vessel(sim, center=250e-6, radius=80e-6), dict(rmax=5e-3)


def mcvox_isopoint_fluencerate(mc, **kw):
    vox = _vox_grid(mc, n=(20, 20, 20))
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='fluence')
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg), mc.mcsource.IsotropicPoint((0, 0, -0.2e-3)),
                detectors=det, fluence=flu, rnginit=97531, **kw)
    return _fill_skin_vessel(sim, center=250e-6, radius=80e-6), dict(rmax=5e-3)


def mcvox_gk2_line_total(mc, **kw):
    vox = _vox_grid(mc, n=(16, 16, 16))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), bottom=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, lambda g: mc.mcpf.Gk2(min(g, 0.9), 0.6, -0.3, 0.5, 0.1)),
                mc.mcsource.Line(), detectors=det, rnginit=8642, **kw)
    return _fill_skin_vessel(sim, center=200e-6, radius=60e-6), dict(rmax=5e-3)


def mcvox_ubeam_radial(mc, **kw):
    """mcvox UniformBeam (tilted, elliptical) launched from above the box."""
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc, n=(20, 20, 16))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(A(0, 0.3e-3, 15)),
                                  bottom=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg),
                mc.mcsource.UniformBeam((120e-6, 80e-6), position=(10e-6, 0, -0.1e-3),
                                        direction=(0.15, -0.1, 1.0)),
                detectors=det, rnginit=2244, **kw)
    return _fill_skin_vessel(sim, center=200e-6, radius=60e-6), dict(rmax=5e-3)


def mcvox_ufiber_fluence(mc, **kw):
    """mcvox UniformFiber in contact with the top surface + deposition grid."""
    vox = _vox_grid(mc, n=(20, 20, 16))
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg),
                mc.mcsource.UniformFiber(_fiber(mc), position=(5e-6, -5e-6, 0.0),
                                         direction=(0.0, 0.1, 1.0)),
                detectors=det, fluence=flu, rnginit=4466, **kw)
    return _fill_skin_vessel(sim, center=200e-6, radius=60e-6), dict(rmax=5e-3)


MCVOX_CASES = {
    'mcvox_ubeam_radial': mcvox_ubeam_radial,
    'mcvox_ufiber_fluence': mcvox_ufiber_fluence,
    'mcvox_gk2_line_total': mcvox_gk2_line_total,
    'mcvox_gauss_fluence': mcvox_gauss_fluence,
    'mcvox_line_mhg_trace': mcvox_line_mhg_trace,
    'mcvox_line_mhg_trace_startend': mcvox_line_mhg_trace_startend,
    'mcvox_isopoint_fluencerate': mcvox_isopoint_fluencerate,
}
ALL_CASES.update(MCVOX_CASES)
GEOMETRY.update({name: 'mcvox' for name in MCVOX_CASES})
GOLDEN_RUN.update({'mcvox_ubeam_radial': (1500, 16), 'mcvox_ufiber_fluence': (1500, 16),
                   'mcvox_gk2_line_total': (1000, 16), 'mcvox_gauss_fluence': (2000, 16), 'mcvox_line_mhg_trace': (600, 16),
                   'mcvox_line_mhg_trace_startend': (1500, 16),
                   'mcvox_isopoint_fluencerate': (2000, 16)})


# ---------------------------------------------------------------------------
# cylindrical cases (layer 0 = surrounding medium, diameters decrease inwards)
def _cyl_layers(mc, pf, stack='two'):
    L = mc.mclayer.Layer
    if stack == 'one':        # BASELINE config 5, mccyl variant: one cylinder r = 5 mm
        return mc.mclayer.Layers([
            L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
            L(d=10e-3, n=1.337, mua=1e2, mus=100e2, pf=pf)])
    return mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=6e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=3e-3, n=1.4, mua=0.5e2, mus=50e2, pf=pf),
        L(d=1e-3, n=1.4, mua=5e2, mus=20e2, pf=pf)])


def mccyl_hg_line_fiz(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        outer=mc.mcdetector.FiZ(Axis(-np.pi, np.pi, 64), Axis(-5e-3, 5e-3, 100)),
        specular=mc.mcdetector.Total())
    return mc.Mc(_cyl_layers(mc, mc.mcpf.Hg(0.8), 'one'),
                 mc.mcsource.Line((-10e-3, 0.0, 0.0), (1.0, 0.0, 0.0)), det,
                 rnginit=RNGINIT, **kw), dict(rmax=25e-3)


def mccyl_mhg_gauss_total(mc, fluence=False, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(outer=mc.mcdetector.Total(cosmin=0.5),
                                  specular=mc.mcdetector.FiZ(Axis(-np.pi, np.pi, 16)))
    flu = mc.mcfluence.FluenceRz(Axis(0, 3e-3, 30), Axis(-2e-3, 2e-3, 40)) if fluence else None
    return mc.Mc(_cyl_layers(mc, mc.mcpf.MHg(0.8, 0.9)),
                 mc.mcsource.GaussianBeam(0.4e-3, position=(-8e-3, 0.5e-3, 0.0),
                                          direction=(1.0, 0.1, 0.05)), det,
                 fluence=flu, rnginit=31337, **kw), dict(rmax=20e-3)


def mccyl_mhg_gauss_total_flurz(mc, **kw):
    return mccyl_mhg_gauss_total(mc, fluence=True, **kw)


def mccyl_gk_ubeam_fiz_trace(mc, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        outer=mc.mcdetector.FiZ(Axis(-np.pi, np.pi, 12), Axis(-3e-3, 3e-3, 10), cosmin=0.2),
        specular=mc.mcdetector.Total())
    tr = mc.mctrace.Trace(maxlen=60, options=mc.mctrace.Trace.TRACE_ALL, plon=True)
    return mc.Mc(_cyl_layers(mc, mc.mcpf.Gk(0.7, 0.5)),
                 mc.mcsource.UniformBeam((1e-3, 5.8e-3), position=(-6e-3, 0.0, 0.0)), det,
                 trace=tr, rnginit=8642, **kw), dict(rmax=15e-3)


def mccyl_hg_isopoint_inside(mc, fluence=False, **kw):
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(outer=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    flu = mc.mcfluence.Fluence(Axis(-3e-3, 3e-3, 12), Axis(-3e-3, 3e-3, 12),
                               Axis(-2e-3, 2e-3, 8), mode='deposition') if fluence else None
    return mc.Mc(_cyl_layers(mc, mc.mcpf.Hg(0.0)),
                 mc.mcsource.IsotropicPoint((1.0e-3, 0.2e-3, 0.0)), det,
                 fluence=flu, rnginit=1122, **kw), dict(rmax=10e-3)


def mccyl_hg_isopoint_fluence(mc, **kw):
    return mccyl_hg_isopoint_inside(mc, fluence=True, **kw)


def mccyl_hg_isopoint_outside(mc, **kw):
    det = mc.mcdetector.Detectors(outer=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    return mc.Mc(_cyl_layers(mc, mc.mcpf.Hg(0.9)),
                 mc.mcsource.IsotropicPoint((-4e-3, 0.0, 0.0)), det,
                 rnginit=2233, **kw), dict(rmax=float('inf'))


MCCYL_CASES = {
    'mccyl_hg_line_fiz': mccyl_hg_line_fiz,
    'mccyl_mhg_gauss_total': mccyl_mhg_gauss_total,
    'mccyl_gk_ubeam_fiz_trace': mccyl_gk_ubeam_fiz_trace,
    'mccyl_hg_isopoint_inside': mccyl_hg_isopoint_inside,
    'mccyl_hg_isopoint_outside': mccyl_hg_isopoint_outside,
}
ALL_CASES.update(MCCYL_CASES)
GEOMETRY.update({name: 'mccyl' for name in MCCYL_CASES})
GOLDEN_RUN.update({'mccyl_hg_line_fiz': (1500, 16), 'mccyl_mhg_gauss_total': (2000, 16),
                   'mccyl_gk_ubeam_fiz_trace': (600, 16), 'mccyl_hg_isopoint_inside': (2000, 16),
                   'mccyl_hg_isopoint_outside': (3000, 16)})

# Configurations the reference cannot build (its mccyl fluence wrapper calls
# mcsim_fluence_deposit with a position argument the header does not declare,
# mccyl.template.c:404-421): no golden vectors exist; the CUDA path is compared
# with the oracle's restatement of the evident semantics only ("parity unpinned").
UNPINNED_CASES = {
    'mccyl_mhg_gauss_total_flurz': mccyl_mhg_gauss_total_flurz,
    'mccyl_hg_isopoint_fluence': mccyl_hg_isopoint_fluence,
}
UNPINNED_GEOMETRY = {name: 'mccyl' for name in UNPINNED_CASES}
UNPINNED_RUN = {name: (2000, 16) for name in UNPINNED_CASES}
# mcpf/pc.py declares `cl_type(mc)` without `self` / @staticmethod: the reference
# cannot pack a Pc layer (TypeError), so Pc is pinned through MPc (same sampling
# branch, reference golden above) and checked oracle <-> GPU only
UNPINNED_CASES['mcml_pf_pc'] = mcml_pf_pc
UNPINNED_GEOMETRY['mcml_pf_pc'] = 'mcml'
UNPINNED_RUN['mcml_pf_pc'] = (3000, 16)


def mcml_hg_isopoint_outside(mc, **kw):
    """Point source above the sample: launch refracts through the top surface
    (the path of mcsource/point.py:96-121 that modifies the position)."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Radial(Axis(0, 5e-3, 50)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.9)),
                 mc.mcsource.IsotropicPoint((0.2e-3, -0.1e-3, -1e-3)), det,
                 rnginit=60606, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_isopoint_outside'] = mcml_hg_isopoint_outside
ALL_CASES['mcml_hg_isopoint_outside'] = mcml_hg_isopoint_outside
GEOMETRY['mcml_hg_isopoint_outside'] = 'mcml'
GOLDEN_RUN['mcml_hg_isopoint_outside'] = (3000, 16)


# ---------------------------------------------------------------------------
# sampling volumes evaluated on the traces of the cases above (config 4)
def make_sv(mc, name):
    A = mc.mcsv.Axis
    return {
        'mcml_lut_iso_radialpl_trace': lambda: mc.mcsv.SamplingVolume(
            A(-1e-3, 1e-3, 20), A(-1e-3, 1e-3, 16), A(0.0, 3e-3, 30)),
        'mcvox_line_mhg_trace': lambda: mc.mcsv.SamplingVolume(
            A(-0.2e-3, 0.2e-3, 16), A(-0.2e-3, 0.2e-3, 16), A(0.0, 0.5e-3, 20)),
        'mccyl_gk_ubeam_fiz_trace': lambda: mc.mcsv.SamplingVolume(
            A(-3e-3, 3e-3, 24), A(-3e-3, 3e-3, 24), A(-1e-3, 1e-3, 8)),
    }[name]()


SV_CASES = ('mcml_lut_iso_radialpl_trace', 'mcvox_line_mhg_trace', 'mccyl_gk_ubeam_fiz_trace')
# binary64: the reference's SamplingVolume kernel rendered in double precision (no C restatement)
SV_DOUBLE_CASES = {'mcml_double_lut_iso_radialpl_trace': 'mcml_lut_iso_radialpl_trace'}


# ---------------------------------------------------------------------------
# user-written plugins (OpenCL-C fragments, tests/user_plugins.py)
def mcml_user_plugins(mc, **kw):
    """Phase function, source and top detector written by a *user* against the
    reference's kernel API; bottom / specular detectors built in."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=up.user_radial(mc, Axis(0, 10e-3, 100)),
        bottom=mc.mcdetector.Radial(Axis(0, 10e-3, 100), cosmin=0.5),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, up.user_hg(mc, 0.8)), up.user_pencil(mc), det,
                 rnginit=424242, **kw), dict(rmax=20e-3)


def mcml_user_plugins_native(mc, **kw):
    """The same simulation with the built-in Hg / Line / Radial plugins: its
    results must equal those of ``mcml_user_plugins`` bit for bit."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(Axis(0, 10e-3, 100)),
        bottom=mc.mcdetector.Radial(Axis(0, 10e-3, 100), cosmin=0.5),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det,
                 rnginit=424242, **kw), dict(rmax=20e-3)


def mcml_user_cubic(mc, **kw):
    """A phase function the reference does not ship (p ~ 1 + cos^2), user-written,
    with a user-written bottom and specular detector around built-in plugins."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(Axis(0, 10e-3, 100)),
        bottom=up.user_radial(mc, Axis(0, 10e-3, 50), cosmin=0.3),
        specular=up.user_radial(mc, Axis(0, 1e-3, 4)))
    return mc.Mc(_layers(mc, up.user_cubic(mc, 1.0)), mc.mcsource.Line(), det,
                 rnginit=515151, **kw), dict(rmax=20e-3)


def mcml_user_fluence(mc, **kw):
    """A fluence accumulator written by a user (1-D deposition profile over depth)
    around built-in plugins."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), bottom=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det,
                 fluence=up.user_depth(mc, Axis(0.0, 3e-3, 60)),
                 rnginit=616161, **kw), dict(rmax=20e-3)


def mcml_rayleigh_line_radial(mc, **kw):
    """Rayleigh phase function (molecular anisotropy 0.3) in a 2-layer stack; the
    reference's class with its (non-compiling) fragment repaired, tests/user_plugins.py."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 10e-3, 100)),
                                  bottom=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, up.rayleigh(mc, 0.3)), mc.mcsource.Line(), det,
                 rnginit=271828, **kw), dict(rmax=20e-3)


def mcml_user_trace(mc, **kw):
    """``mcml_lut_iso_radialpl_trace`` with a user-written trace that restates the built-in
    event record: equal to it bit for bit."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    params, lut = _hg_lut()
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.RadialPl(Axis(0, 5e-3, 50), Axis(0, 0.05, 100)),
        bottom=mc.mcdetector.TotalPl(Axis(0, 0.05, 100)))
    tr = up.user_trace(mc, maxlen=50, options=mc.mctrace.Trace.TRACE_ALL)
    return mc.Mc(_layers(mc, mc.mcpf.Lut(params, lut)),
                 mc.mcsource.IsotropicPoint((0, 0, 0.5e-3)), det, trace=tr,
                 rnginit=777, **kw), dict(rmax=20e-3)


def mcml_user_trace_squared(mc, **kw):
    """A user-written trace that records weight^2 (not in the reference), start / end
    events only, event mask on the boundary events."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), bottom=mc.mcdetector.Total())
    T = mc.mctrace.Trace
    tr = up.user_trace(mc, squared=True, maxlen=40, options=T.TRACE_ALL,
                       event_mask=T.TRACE_EVENT_BOUNDARY_HIT | T.TRACE_EVENT_LAUNCH)
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line((0, 0, 0), (0.2, 0, 1)), det,
                 trace=tr, rnginit=818181, **kw), dict(rmax=20e-3)


def _with_user_trace(native_case, **trace_kw):
    """``native_case`` with its Trace swapped for the user-written restatement of the
    built-in event record (tests/user_plugins.py): equal to it bit for bit."""
    def make(mc, **kw):
        import user_plugins as up
        sim, attrs = native_case(mc, **kw)
        old = sim.trace
        sim._trace = up.user_trace(mc, maxlen=old.maxlen, options=old.options, plon=old.plon,
                                   **trace_kw)
        return sim, attrs
    make.__doc__ = _with_user_trace.__doc__
    return make


mcvox_user_trace = _with_user_trace(mcvox_line_mhg_trace)
mccyl_user_trace = _with_user_trace(mccyl_gk_ubeam_fiz_trace)


def mcvox_user_plugins(mc, **kw):
    """Voxel geometry with a user-written source (starts inside the box, uses the voxel /
    material accessors), a user-written top detector and a user-written phase function."""
    import user_plugins as up
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc)
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(
        top=up.user_radial(mc, A(0, 0.4e-3, 40)), bottom=mc.mcdetector.Total())
    src = up.user_vox_beam(mc, (20e-6, -10e-6, 60e-6), (0.3, 0.1, 1.0))
    sim = mc.Mc(vox, _vox_materials(mc, lambda g: up.user_hg(mc, g)), src,
                detectors=det, fluence=flu, rnginit=565656, **kw)
    return _fill_skin_vessel(sim), dict(rmax=25e-3)


def mccyl_user_plugins(mc, **kw):
    """Cylindrical geometry with a user-written source (starts in the second cylinder), a
    user-written outer detector and a user-written phase function."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(outer=up.user_radial(mc, Axis(4e-3, 6e-3, 20), cosmin=0.0),
                                  specular=mc.mcdetector.Total())
    src = up.user_cyl_beam(mc, (1.2e-3, 0.4e-3, 0.1e-3), (0.6, 0.2, 0.3))
    return mc.Mc(_cyl_layers(mc, up.user_hg(mc, 0.8)), src, det,
                 rnginit=575757, **kw), dict(rmax=25e-3)


def mcml_user_surface_reflector(mc, **kw):
    """A top surface layout written by a user (the arithmetic of LambertianReflector):
    equals ``mcml_surface_lambert_top`` bit for bit."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    surf = mc.mcsurface.SurfaceLayouts(top=up.user_reflector(mc, 0.8, 0.0))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(),
                                  bottom=mc.mcdetector.Radial(Axis(0, 5e-3, 50)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det, surface=surf,
                 rnginit=13579, **kw), dict(rmax=20e-3)


def mcml_user_surface_window(mc, **kw):
    """A layout the reference does not ship, on both surfaces: anti-reflection window
    (handler returns MC_REFRACTED), black ring (MC_REFLECTED with zero weight), glass
    elsewhere (n2 / cc override, MC_SURFACE_LAYOUT_CONTINUE)."""
    import user_plugins as up
    Axis = mc.mcdetector.Axis
    surf = mc.mcsurface.SurfaceLayouts(top=up.user_window(mc, 0.4e-3, 0.8e-3, 1.52),
                                       bottom=up.user_window(mc, 1.0e-3, 1.5e-3, 1.45))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 4e-3, 80)),
                                  bottom=mc.mcdetector.Radial(Axis(0, 4e-3, 80)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det, surface=surf,
                 rnginit=717171, **kw), dict(rmax=20e-3)


for _name, _make in (('mcml_rayleigh_line_radial', mcml_rayleigh_line_radial),):
    MCML_CASES[_name] = ALL_CASES[_name] = _make
    GEOMETRY[_name] = 'mcml'
    GOLDEN_RUN[_name] = (3000, 16)
MCML_CASES['mcml_user_plugins_native'] = mcml_user_plugins_native
ALL_CASES['mcml_user_plugins_native'] = mcml_user_plugins_native
GEOMETRY['mcml_user_plugins_native'] = 'mcml'
GOLDEN_RUN['mcml_user_plugins_native'] = (3000, 16)
# cases with user fragments: golden vectors come from the reference kernel
# executing the same fragments; there is no C restatement of user code, so the
# oracle pins them through the equivalent built-in case (value) where one exists
USER_CASES = {'mcml_user_plugins': mcml_user_plugins, 'mcml_user_cubic': mcml_user_cubic,
              'mcml_user_fluence': mcml_user_fluence,
              'mcml_user_surface_reflector': mcml_user_surface_reflector,
              'mcml_user_surface_window': mcml_user_surface_window,
              'mcvox_user_plugins': mcvox_user_plugins,
              'mccyl_user_plugins': mccyl_user_plugins,
              'mcml_user_trace': mcml_user_trace,
              'mcml_user_trace_squared': mcml_user_trace_squared,
              'mcvox_user_trace': mcvox_user_trace, 'mccyl_user_trace': mccyl_user_trace}
USER_EQUIVALENT = {'mcml_user_plugins': 'mcml_user_plugins_native', 'mcml_user_cubic': None,
                   'mcml_user_fluence': None,
                   'mcml_user_surface_reflector': 'mcml_surface_lambert_top',
                   'mcml_user_surface_window': None,
                   'mcvox_user_plugins': None,
                   'mccyl_user_plugins': None,
                   'mcml_user_trace': 'mcml_lut_iso_radialpl_trace',
                   'mcml_user_trace_squared': None,
                   'mcvox_user_trace': 'mcvox_line_mhg_trace',
                   'mccyl_user_trace': 'mccyl_gk_ubeam_fiz_trace'}
USER_GEOMETRY = {name: name.split('_')[0] for name in USER_CASES}
USER_RUN = {name: (3000, 16) for name in USER_CASES}
USER_RUN['mcml_user_trace'] = (800, 16)
USER_RUN['mcml_user_trace_squared'] = (800, 16)
USER_RUN['mcvox_user_trace'] = (600, 16)
USER_RUN['mccyl_user_trace'] = (600, 16)


# ---------------------------------------------------------------------------
# fiber-array probes (mcdetector/probe/lineararray.py, fiberarray.py)
def _fiber_layout(mc, fib, position, direction=(0.0, 0.0, 1.0)):
    if mc.__name__.startswith('xopto'):
        from xopto.mcml.mcutil import fiber as fiberutil
        return fiberutil.FiberLayout(fib, position, direction)
    return mc.mcdetector.FiberLayout(fib, position, direction)


def mcml_hg_fiber_lineararray(mc, **kw):
    """UniformFiber source under a tilted 5-fiber linear array (top) and a
    3-fiber array along y (bottom)."""
    fib = _fiber(mc)
    tilt = (np.sin(np.deg2rad(8.0)), 0.0, np.cos(np.deg2rad(8.0)))
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.LinearArray(fib, 5, spacing=250e-6, direction=tilt,
                                      position=(0.1e-3, 0.0)),
        bottom=mc.mcdetector.LinearArray(fib, 3, orientation=(0.0, 1.0)),
        specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.UniformFiber(fib), det,
                 rnginit=737373, **kw), dict(rmax=20e-3)


def mcml_mhg_line_fiberarray(mc, **kw):
    """Line source, individually placed and tilted fibers of two kinds (top),
    a two-fiber array at the bottom."""
    fib = _fiber(mc)
    if mc.__name__.startswith('xopto'):
        from xopto.mcml.mcutil import fiber as fiberutil
        big = fiberutil.MultimodeFiber(400e-6, 440e-6, 1.462, 0.37)
    else:
        big = mc.mcsource.MultimodeFiber(400e-6, 440e-6, 1.462, 0.37)
    tilt = (0.0, np.sin(np.deg2rad(-10.0)), np.cos(np.deg2rad(-10.0)))
    top = mc.mcdetector.FiberArray([
        _fiber_layout(mc, fib, (0.0, 0.0, 0.0)),
        _fiber_layout(mc, big, (0.5e-3, 0.0, 0.0), tilt),
        _fiber_layout(mc, fib, (-0.3e-3, 0.3e-3, 0.0)),
        _fiber_layout(mc, big, (0.0, -0.6e-3, 0.0))])
    bottom = mc.mcdetector.FiberArray([
        _fiber_layout(mc, big, (0.0, 0.0, 0.0)), _fiber_layout(mc, big, (0.6e-3, 0.0, 0.0))])
    det = mc.mcdetector.Detectors(top=top, bottom=bottom)
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.7, 0.8)), mc.mcsource.Line(), det,
                 rnginit=848484, **kw), dict(rmax=20e-3)


for _name, _fn in (('mcml_hg_fiber_lineararray', mcml_hg_fiber_lineararray),
                   ('mcml_mhg_line_fiberarray', mcml_mhg_line_fiberarray)):
    MCML_CASES[_name] = _fn
    ALL_CASES[_name] = _fn
    GEOMETRY[_name] = 'mcml'
    GOLDEN_RUN[_name] = (4000, 16)


def mcml_hg_fiber_arrays_pl(mc, **kw):
    """Path-length resolved fiber arrays: LinearArrayPl (top, log path-length
    axis) and FiberArrayPl (bottom)."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    top = mc.mcdetector.LinearArrayPl(fib, 4, spacing=300e-6,
                                      plaxis=Axis(1e-4, 1e-1, 12, logscale=True))
    bottom = mc.mcdetector.FiberArrayPl(
        [_fiber_layout(mc, fib, (0.0, 0.0, 0.0)), _fiber_layout(mc, fib, (0.4e-3, 0.1e-3, 0.0)),
         _fiber_layout(mc, fib, (-0.4e-3, 0.0, 0.0))], plaxis=Axis(0.0, 30e-3, 15))
    det = mc.mcdetector.Detectors(top=top, bottom=bottom, specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.9)), mc.mcsource.UniformFiber(fib), det,
                 rnginit=959595, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_fiber_arrays_pl'] = mcml_hg_fiber_arrays_pl
ALL_CASES['mcml_hg_fiber_arrays_pl'] = mcml_hg_fiber_arrays_pl
GEOMETRY['mcml_hg_fiber_arrays_pl'] = 'mcml'
GOLDEN_RUN['mcml_hg_fiber_arrays_pl'] = (6000, 16)


def mcml_hg_fiber_fluencecylt(mc, **kw):
    """Time-resolved cylindrical deposition grid (FluenceCylt)."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceCylt(Axis(0, 2e-3, 10), Axis(0, 2*np.pi, 6), Axis(0, 3e-3, 15),
                                   Axis(0, 40e-12, 8), center=(0.05e-3, -0.05e-3))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.UniformFiber(fib, position=(0.1e-3, 0, 0), direction=(0.2, 0.1, 1)),
                 det, fluence=flu, rnginit=80808, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_fiber_fluencecylt'] = mcml_hg_fiber_fluencecylt
ALL_CASES['mcml_hg_fiber_fluencecylt'] = mcml_hg_fiber_fluencecylt
GEOMETRY['mcml_hg_fiber_fluencecylt'] = 'mcml'
GOLDEN_RUN['mcml_hg_fiber_fluencecylt'] = (2000, 16)


def mcml_mhg_lambertianfiber_radial(mc, **kw):
    """Tilted LambertianFiber source (mcsource/fiber.py:499), radial detectors."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.8, 0.7)),
                 mc.mcsource.LambertianFiber(fib, position=(0.2e-3, -0.1e-3, 0),
                                             direction=(0.1, -0.2, 1)),
                 det, rnginit=171717, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_mhg_lambertianfiber_radial'] = mcml_mhg_lambertianfiber_radial
ALL_CASES['mcml_mhg_lambertianfiber_radial'] = mcml_mhg_lambertianfiber_radial
GEOMETRY['mcml_mhg_lambertianfiber_radial'] = 'mcml'
GOLDEN_RUN['mcml_mhg_lambertianfiber_radial'] = (3000, 16)


def mcvox_isovoxel_fluence(mc, **kw):
    """IsotropicVoxel source inside the vessel (mcvox/mcsource/voxel.py), deposition
    grid and radial / total detectors."""
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc)
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(A(0, 0.4e-3, 20)), bottom=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg), mc.mcsource.IsotropicVoxel((13, 9, 12)),
                detectors=det, fluence=flu, rnginit=363636, **kw)
    return _fill_skin_vessel(sim), dict(rmax=25e-3)


MCVOX_CASES['mcvox_isovoxel_fluence'] = mcvox_isovoxel_fluence
ALL_CASES['mcvox_isovoxel_fluence'] = mcvox_isovoxel_fluence
GEOMETRY['mcvox_isovoxel_fluence'] = 'mcvox'
GOLDEN_RUN['mcvox_isovoxel_fluence'] = (2000, 16)


def mcml_mhg_gauss_enhanced_rng(mc, **kw):
    """MC_USE_ENHANCED_RNG: two MWC steps (64 random bits) per uniform draw."""
    Axis = mc.mcdetector.Axis
    opts = list(kw.pop('options', None) or []) + [mc.mcoptions.McUseEnhancedRng.on]
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 20), Axis(0, 3e-3, 30))
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.8, 0.9)), mc.mcsource.GaussianBeam(100e-6), det,
                 fluence=flu, rnginit=191919, options=opts, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_mhg_gauss_enhanced_rng'] = mcml_mhg_gauss_enhanced_rng
ALL_CASES['mcml_mhg_gauss_enhanced_rng'] = mcml_mhg_gauss_enhanced_rng
GEOMETRY['mcml_mhg_gauss_enhanced_rng'] = 'mcml'
GOLDEN_RUN['mcml_mhg_gauss_enhanced_rng'] = (2000, 16)


def mcml_surface_lineararray(mc, **kw):
    """Linear fiber-array probe on both surfaces (mcml/mcsurface/probe/lineararray.py):
    cores / claddings, a rectangular filled cut-out and the reflective tip on top
    (tilted, rotated array), a plain array without cut-out at the bottom; the
    matching LinearArray detectors collect the light."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    L = mc.mclayer.Layer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=0.3e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=0.4e-3, n=1.4, mua=0.5e2, mus=50e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])
    orient = (np.cos(np.deg2rad(30.0)), np.sin(np.deg2rad(30.0)))
    surf = mc.mcsurface.SurfaceLayouts(
        top=mc.mcsurface.LinearArray(fib, 3, spacing=260e-6, orientation=orient, diameter=2.5e-3,
                                     reflectivity=0.55, cutout=(1.0e-3, 0.4e-3), cutoutn=1.6,
                                     position=(30e-6, -20e-6), direction=(0.04, 0.0, 1.0)),
        bottom=mc.mcsurface.LinearArray(fib, 2, diameter=1.5e-3, reflectivity=0.8))
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.LinearArray(fib, 3, spacing=260e-6, orientation=orient,
                                      position=(30e-6, -20e-6), direction=(0.04, 0.0, 1.0)),
        bottom=mc.mcdetector.LinearArray(fib, 2), specular=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 40), Axis(0, 0.7e-3, 35))
    return mc.Mc(layers, mc.mcsource.UniformFiber(fib), det, fluence=flu, surface=surf,
                 rnginit=13579, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_surface_lineararray'] = mcml_surface_lineararray
ALL_CASES['mcml_surface_lineararray'] = mcml_surface_lineararray
GEOMETRY['mcml_surface_lineararray'] = 'mcml'
GOLDEN_RUN['mcml_surface_lineararray'] = (4000, 16)


def mcml_surface_fiberarray(mc, **kw):
    """FiberArray probe layout on top (two fiber kinds, one tilted), matching
    FiberArray detector; the tip reflectivity never reaches the reference kernel
    (see pyxopto_b200.mcml.mcsurface.FiberArray), so the tip absorbs."""
    Axis = mc.mcdetector.Axis
    fib = _fiber(mc)
    if mc.__name__.startswith('xopto'):
        from xopto.mcml.mcutil import fiber as fiberutil
        big = fiberutil.MultimodeFiber(400e-6, 440e-6, 1.462, 0.37)
    else:
        big = mc.mcsource.MultimodeFiber(400e-6, 440e-6, 1.462, 0.37)
    L = mc.mclayer.Layer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=0.3e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=0.4e-3, n=1.4, mua=0.5e2, mus=50e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])
    tilt = (0.0, np.sin(np.deg2rad(-10.0)), np.cos(np.deg2rad(-10.0)))
    fibers = [_fiber_layout(mc, fib, (0.0, 0.0, 0.0)),
              _fiber_layout(mc, big, (0.5e-3, 0.0, 0.0), tilt),
              _fiber_layout(mc, fib, (-0.3e-3, 0.3e-3, 0.0))]
    surf = mc.mcsurface.SurfaceLayouts(
        top=mc.mcsurface.FiberArray(fibers, diameter=2e-3, reflectivity=0.7,
                                    position=(0.1e-3, 0.05e-3)))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.FiberArray(fibers),
                                  bottom=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    return mc.Mc(layers, mc.mcsource.UniformFiber(fib), det, surface=surf,
                 rnginit=97531, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_surface_fiberarray'] = mcml_surface_fiberarray
ALL_CASES['mcml_surface_fiberarray'] = mcml_surface_fiberarray
GEOMETRY['mcml_surface_fiberarray'] = 'mcml'
GOLDEN_RUN['mcml_surface_fiberarray'] = (4000, 16)


def mcml_hg_line_totallut(mc, **kw):
    """TotalLut detectors: angular sensitivity tables in the float lookup-table pool
    (mcdetector/total.py:240, mcutil/lut.py CollectionLut).  The reference's
    TotalLut does not switch MC_USE_FP_LUT on itself, so its kernel only builds
    next to another table user - here the Lut phase function."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import CollectionLut
    else:
        CollectionLut = mc.mcdetector.CollectionLut
    ct = np.linspace(0.0, 1.0, 21)
    top = mc.mcdetector.TotalLut(CollectionLut(ct**2, ct, n=100))
    bottom = mc.mcdetector.TotalLutPl(CollectionLut(np.sqrt(np.linspace(0.2, 1.0, 9)),
                                                    np.linspace(0.2, 1.0, 9), n=33),
                                      plaxis=mc.mcdetector.Axis(0.0, 40e-3, 800),
                                      direction=(0.1, 0.0, 1.0))
    det = mc.mcdetector.Detectors(top=top, bottom=bottom, specular=mc.mcdetector.Total())
    params, lut = _hg_lut()
    return mc.Mc(_layers(mc, mc.mcpf.Lut(params, lut)), mc.mcsource.Line(), det,
                 rnginit=262626, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_line_totallut'] = mcml_hg_line_totallut
ALL_CASES['mcml_hg_line_totallut'] = mcml_hg_line_totallut
GEOMETRY['mcml_hg_line_totallut'] = 'mcml'
GOLDEN_RUN['mcml_hg_line_totallut'] = (3000, 16)


def mcml_lut_ufiberlut_totallut(mc, **kw):
    """UniformFiberLut source (tabulated emission, refracted into the sample) under a
    tilted TotalLut detector (mcsource/fiber.py:690, mcutil/lut.py EmissionLut)."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import CollectionLut, EmissionLut
        from xopto.mcml.mcutil.fiber import MultimodeFiberLut
    else:
        CollectionLut, EmissionLut = mc.mcdetector.CollectionLut, mc.mcsource.EmissionLut
        MultimodeFiberLut = mc.mcsource.MultimodeFiberLut
    ct = np.linspace(np.cos(np.deg2rad(25.0)), 1.0, 40)
    emission = EmissionLut(np.exp(-((1.0 - ct)/0.03)**2), ct, n=200, npts=2000)
    fib = MultimodeFiberLut(200e-6, 220e-6, 1.462, None, emission=emission)
    cs = np.linspace(0.0, 1.0, 11)
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.TotalLut(CollectionLut(cs, cs, n=64), direction=(0.0, 0.1, 1.0)),
        bottom=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    params, lut = _hg_lut()
    return mc.Mc(_layers(mc, mc.mcpf.Lut(params, lut)),
                 mc.mcsource.UniformFiberLut(fib, position=(0.1e-3, 0, 0), direction=(0.15, 0.0, 1)),
                 det, rnginit=383838, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_lut_ufiberlut_totallut'] = mcml_lut_ufiberlut_totallut
ALL_CASES['mcml_lut_ufiberlut_totallut'] = mcml_lut_ufiberlut_totallut
GEOMETRY['mcml_lut_ufiberlut_totallut'] = 'mcml'
GOLDEN_RUN['mcml_lut_ufiberlut_totallut'] = (3000, 16)


def mcml_hg_line_fiberlutarray(mc, **kw):
    """FiberLutArray detector: fibers with tabulated collection sensitivity
    (mcdetector/probe/fiberlutarray.py)."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import CollectionLut
        from xopto.mcml.mcutil.fiber import MultimodeFiberLut
    else:
        CollectionLut = mc.mcdetector.CollectionLut
        MultimodeFiberLut = mc.mcsource.MultimodeFiberLut
    cs = np.linspace(0.5, 1.0, 26)
    fa = MultimodeFiberLut(400e-6, 440e-6, 1.462, None,
                           collection=CollectionLut((cs - 0.5)*2.0, cs, n=50))
    fb = MultimodeFiberLut(600e-6, 660e-6, 1.462, None,
                           collection=CollectionLut(np.sqrt(cs), cs, n=80))
    tilt = (np.sin(np.deg2rad(6.0)), 0.0, np.cos(np.deg2rad(6.0)))
    top = mc.mcdetector.FiberLutArray([
        _fiber_layout(mc, fa, (0.3e-3, 0.0, 0.0)),
        _fiber_layout(mc, fb, (-0.5e-3, 0.2e-3, 0.0), tilt),
        _fiber_layout(mc, fa, (0.0, -0.6e-3, 0.0))])
    det = mc.mcdetector.Detectors(top=top, bottom=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)), mc.mcsource.Line(), det,
                 rnginit=484848, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_line_fiberlutarray'] = mcml_hg_line_fiberlutarray
ALL_CASES['mcml_hg_line_fiberlutarray'] = mcml_hg_line_fiberlutarray
GEOMETRY['mcml_hg_line_fiberlutarray'] = 'mcml'
GOLDEN_RUN['mcml_hg_line_fiberlutarray'] = (4000, 16)


def mcml_mhg_rect_uniform(mc, **kw):
    """UniformRectangular source on the surface (mcsource/rectangular.py:32)."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Cartesian(Axis(-2e-3, 2e-3, 20)),
                                  bottom=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.8, 0.9)),
                 mc.mcsource.UniformRectangular(1e-3, 0.5e-3, 1.45, 0.4, position=(0.1e-3, 0, 0)),
                 det, rnginit=575757, **kw), dict(rmax=20e-3)


def mcml_hg_rect_lambertian_inside(mc, **kw):
    """LambertianRectangular source buried in the second sample layer."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 20), Axis(0, 3e-3, 30))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.LambertianRectangular(0.4e-3, 0.8e-3, 1.6, 0.5,
                                                   position=(0, -0.1e-3, 1.5e-3)),
                 det, fluence=flu, rnginit=676767, **kw), dict(rmax=20e-3)


for _name, _fn in (('mcml_mhg_rect_uniform', mcml_mhg_rect_uniform),
                   ('mcml_hg_rect_lambertian_inside', mcml_hg_rect_lambertian_inside)):
    MCML_CASES[_name] = _fn
    ALL_CASES[_name] = _fn
    GEOMETRY[_name] = 'mcml'
    GOLDEN_RUN[_name] = (3000, 16)


def mcvox_isovoxels_total(mc, **kw):
    """IsotropicVoxels source: a weighted set of emitting voxels kept in the float
    lookup-table pool (mcvox/mcsource/voxel.py:195)."""
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc)
    voxels = np.array([[12, 10, 14], [13, 10, 14], [12, 11, 13], [5, 4, 3], [20, 15, 25]])
    weights = np.array([1.0, 0.5, 0.25, 0.8, 0.1])
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Radial(A(0, 0.4e-3, 20)), bottom=mc.mcdetector.Total())
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg), mc.mcsource.IsotropicVoxels(voxels, weights),
                detectors=det, fluence=flu, rnginit=787878, **kw)
    return _fill_skin_vessel(sim), dict(rmax=25e-3)


MCVOX_CASES['mcvox_isovoxels_total'] = mcvox_isovoxels_total
ALL_CASES['mcvox_isovoxels_total'] = mcvox_isovoxels_total
GEOMETRY['mcvox_isovoxels_total'] = 'mcvox'
GOLDEN_RUN['mcvox_isovoxels_total'] = (2000, 16)


def mcml_hgdir_line_radial(mc, **kw):
    """HgDir: a phase function that samples the new direction itself
    (MC_PF_SAMPLE_DIRECTION, mcpf/hgdir.py)."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 100)),
                                  bottom=mc.mcdetector.Cartesian(Axis(-3e-3, 3e-3, 30)),
                                  specular=mc.mcdetector.Total())
    pf = mc.mcpf.HgDir(0.8, direction=(0.6, 0.0, 0.8), p=0.3)
    return mc.Mc(_layers(mc, pf), mc.mcsource.Line(), det,
                 rnginit=898989, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hgdir_line_radial'] = mcml_hgdir_line_radial
ALL_CASES['mcml_hgdir_line_radial'] = mcml_hgdir_line_radial
GEOMETRY['mcml_hgdir_line_radial'] = 'mcml'
GOLDEN_RUN['mcml_hgdir_line_radial'] = (3000, 16)


def mcml_hg_ufiberni_radial(mc, **kw):
    """UniformFiberNI source (normal incidence, mcsource/fiberni.py:180) with a specular
    detector that sees the direction refracted back into the core."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 100)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Radial(Axis(0, 0.2e-3, 20), cosmin=0.98))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.UniformFiberNI(_fiber(mc), position=(0.05e-3, -0.02e-3, 0)),
                 det, rnginit=919191, **kw), dict(rmax=20e-3)


def mcml_mhg_lfiberni_cart(mc, **kw):
    """LambertianFiberNI source (mcsource/fiberni.py:422)."""
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Cartesian(Axis(-2e-3, 2e-3, 20)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Total())
    flu = mc.mcfluence.FluenceRz(Axis(0, 2e-3, 20), Axis(0, 3e-3, 30))
    return mc.Mc(_layers(mc, mc.mcpf.MHg(0.8, 0.9)),
                 mc.mcsource.LambertianFiberNI(_fiber(mc), position=(-0.1e-3, 0, 0)),
                 det, fluence=flu, rnginit=929292, **kw), dict(rmax=20e-3)


def mcml_hg_ufiberlutni_total(mc, **kw):
    """UniformFiberLutNI source: tabulated emission at normal incidence
    (mcsource/fiberni.py:581)."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import EmissionLut
        from xopto.mcml.mcutil.fiber import MultimodeFiberLut
    else:
        EmissionLut = mc.mcsource.EmissionLut
        MultimodeFiberLut = mc.mcsource.MultimodeFiberLut
    ct = np.linspace(np.cos(np.deg2rad(20.0)), 1.0, 30)
    emission = EmissionLut(np.exp(-((1.0 - ct)/0.02)**2), ct, n=150, npts=2000)
    fib = MultimodeFiberLut(200e-6, 220e-6, 1.462, None, emission=emission)
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Total(),
                                  specular=mc.mcdetector.Total())
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.UniformFiberLutNI(fib, position=(0, 0.1e-3, 0)),
                 det, rnginit=939393, **kw), dict(rmax=20e-3)


for _name, _fn in (('mcml_hg_ufiberni_radial', mcml_hg_ufiberni_radial),
                   ('mcml_mhg_lfiberni_cart', mcml_mhg_lfiberni_cart),
                   ('mcml_hg_ufiberlutni_total', mcml_hg_ufiberlutni_total)):
    MCML_CASES[_name] = _fn
    ALL_CASES[_name] = _fn
    GEOMETRY[_name] = 'mcml'
    GOLDEN_RUN[_name] = (3000, 16)


def mcml_hg_rectlut_inside(mc, **kw):
    """UniformRectangularLut source buried in the second layer: tabulated emission
    (mcsource/rectangular.py:530)."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import EmissionLut
    else:
        EmissionLut = mc.mcsource.EmissionLut
    ct = np.linspace(np.cos(np.deg2rad(40.0)), 1.0, 50)
    emission = EmissionLut(ct**2, ct, n=300, npts=2000)
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0, 5e-3, 50)),
                                  bottom=mc.mcdetector.Cartesian(Axis(-2e-3, 2e-3, 10)))
    return mc.Mc(_layers(mc, mc.mcpf.Hg(0.8)),
                 mc.mcsource.UniformRectangularLut(emission, 0.6e-3, 0.3e-3, 1.55,
                                                   position=(0.1e-3, 0, 1.2e-3)),
                 det, rnginit=949494, **kw), dict(rmax=20e-3)


MCML_CASES['mcml_hg_rectlut_inside'] = mcml_hg_rectlut_inside
ALL_CASES['mcml_hg_rectlut_inside'] = mcml_hg_rectlut_inside
GEOMETRY['mcml_hg_rectlut_inside'] = 'mcml'
GOLDEN_RUN['mcml_hg_rectlut_inside'] = (3000, 16)


def mcvox_lfiber_radial(mc, **kw):
    """mcvox LambertianFiber tilted against the top surface (mcvox/mcsource/fiber.py:523)."""
    A = mc.mcgeometry.Axis
    vox = _vox_grid(mc, n=(20, 20, 16))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(A(0, 0.4e-3, 20)),
                                  bottom=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg),
                mc.mcsource.LambertianFiber(_fiber(mc), position=(-5e-6, 10e-6, 0.0),
                                            direction=(0.1, 0.0, 1.0)),
                detectors=det, rnginit=959595, **kw)
    return _fill_skin_vessel(sim, center=200e-6, radius=60e-6), dict(rmax=5e-3)


def mcvox_ufiberlut_fluence(mc, **kw):
    """mcvox UniformFiberLut: tabulated emission + deposition grid
    (mcvox/mcsource/fiber.py:747)."""
    if mc.__name__.startswith('xopto'):
        from xopto.mcbase.mcutil.lut import EmissionLut
        from xopto.mcvox.mcutil.fiber import MultimodeFiberLut
    else:
        EmissionLut = mc.mcsource.EmissionLut
        MultimodeFiberLut = mc.mcsource.MultimodeFiberLut
    ct = np.linspace(np.cos(np.deg2rad(25.0)), 1.0, 40)
    emission = EmissionLut(np.exp(-((1.0 - ct)/0.03)**2), ct, n=200, npts=2000)
    fib = MultimodeFiberLut(200e-6, 220e-6, 1.462, None, emission=emission)
    vox = _vox_grid(mc, n=(20, 20, 16))
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), specular=mc.mcdetector.Total())
    sim = mc.Mc(vox, _vox_materials(mc, mc.mcpf.Hg),
                mc.mcsource.UniformFiberLut(fib, position=(5e-6, -5e-6, 0.0),
                                            direction=(0.0, 0.1, 1.0)),
                detectors=det, fluence=flu, rnginit=969696, **kw)
    return _fill_skin_vessel(sim, center=200e-6, radius=60e-6), dict(rmax=5e-3)


for _name, _fn in (('mcvox_lfiber_radial', mcvox_lfiber_radial),
                   ('mcvox_ufiberlut_fluence', mcvox_ufiberlut_fluence)):
    MCVOX_CASES[_name] = _fn
    ALL_CASES[_name] = _fn
    GEOMETRY[_name] = 'mcvox'
    GOLDEN_RUN[_name] = (1500, 16)


# ---- the bench configurations (benchcfg.py) at their real size -----------------
# BASELINE.json's configs themselves under the oracle: C2's 5-entry stack with the
# 250 x 500 FluenceRz grid, C3 at 201^3 (compact uint8 map / power-of-two strides /
# 4 GB-window addressing of the CUDA path), C4's maxlen-512 trace, the C5 points.
# (packets, work-items) of the golden run with the reference kernel:
BENCH_RUN = {'c1_slab': (2000, 16), 'c2_skin': (3000, 16), 'c3_vox': (1000, 16),
             'c4_trace': (96, 16), 'c5_cyl': (1000, 16), 'c5_slab': (1000, 16)}


def bench_case(name):
    """Case function ``f(mc, **kw) -> (sim, attrs)`` of bench configuration ``name``."""
    import benchcfg

    def make(mc, **kw):
        return benchcfg.CONFIGS[name](mc, **kw), {}
    return make


def bench_geometry(name):
    import benchcfg
    return benchcfg.GEOMETRY[name]


# ---- per-packet trajectories against the reference kernel -----------------------
# packets = work-items (one packet per MWC stream) of tests/golden/traj_<case>.npz
TRAJ_RUN = {'mcml_lut_iso_radialpl_trace': 512, 'mcvox_line_mhg_trace': 512,
            'mccyl_gk_ubeam_fiz_trace': 512, 'c4_trace': 192}


# ---------------------------------------------------------------------------
# anisotropic layers (mcml/mclayer/layer.py:391-790): absorption / scattering tensors
# projected on the propagation direction
def mcml_aniso_line_cart_flu(mc, **kw):
    Axis = mc.mcdetector.Axis
    L = mc.mclayer.AnisotropicLayer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=1e-3, n=1.33, mua=[1e2, 2e2, 0.5e2],
          mus=np.array([[100e2, 10e2, 0.0], [10e2, 60e2, 5e2], [0.0, 5e2, 150e2]]), pf=pf),
        L(d=2e-3, n=1.4, mua=0.5e2, mus=[50e2, 80e2, 30e2], pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf)])
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.Cartesian(Axis(-2e-3, 2e-3, 40)), bottom=mc.mcdetector.Total(),
        specular=mc.mcdetector.Total())
    flu = mc.mcfluence.Fluence(Axis(-1e-3, 1e-3, 20), Axis(-1e-3, 1e-3, 20),
                               Axis(0, 3e-3, 30), mode='deposition')
    return mc.Mc(layers, mc.mcsource.Line((0.0, 0.0, 0.0), (0.2, 0.1, 1.0)), det,
                 fluence=flu, rnginit=99, **kw), dict(rmax=20e-3)


ALL_CASES['mcml_aniso_line_cart_flu'] = mcml_aniso_line_cart_flu
GEOMETRY['mcml_aniso_line_cart_flu'] = 'mcml'
GOLDEN_RUN['mcml_aniso_line_cart_flu'] = (2000, 16)


def mcvox_aniso_gauss_fluence(mc, **kw):
    """AnisotropicMaterial (mcbase/mcmaterial.py:310-655) in the voxel geometry."""
    A = mc.mcgeometry.Axis
    vox = mc.mcgeometry.Voxels(A(-300e-6, 300e-6, 24), A(-250e-6, 250e-6, 20), A(0.0, 700e-6, 28))
    M = mc.mcmaterial.AnisotropicMaterial
    pf = mc.mcpf.Hg(0.8)
    mats = mc.mcmaterial.Materials([
        M(n=1.0, mua=0.0, mus=0.0, pf=pf),
        M(n=1.33, mua=[20e2, 10e2, 5e2],
          mus=np.array([[300e2, 20e2, 0.0], [20e2, 200e2, 10e2], [0.0, 10e2, 400e2]]), pf=pf),
        M(n=1.4, mua=8e2, mus=[150e2, 250e2, 100e2], pf=pf)])
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Total(), bottom=mc.mcdetector.Total())
    sim = mc.Mc(vox, mats, mc.mcsource.GaussianBeam(40e-6), det, fluence=flu,
                rnginit=4711, **kw)
    z, y, x = sim.voxels.meshgrid()
    m = sim.voxels.material
    m[z <= 200e-6] = 1
    m[z > 200e-6] = 2
    m[(x**2 + (z - 350e-6)**2) <= (90e-6)**2] = 1
    return sim, dict(rmax=5e-3)


def mccyl_aniso_line_fiz(mc, **kw):
    """AnisotropicLayer (mccyl/mclayer/layer.py:455-760) in the cylindrical geometry."""
    Axis = mc.mcdetector.Axis
    # (the reference's mccyl.mclayer package does not re-export the class)
    L = getattr(mc.mclayer, 'AnisotropicLayer', None) or mc.mclayer.layer.AnisotropicLayer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=6e-3, n=1.33, mua=[1e2, 2e2, 0.5e2],
          mus=np.array([[100e2, 10e2, 0.0], [10e2, 60e2, 5e2], [0.0, 5e2, 150e2]]), pf=pf),
        L(d=3e-3, n=1.4, mua=0.5e2, mus=[50e2, 80e2, 30e2], pf=pf)])
    det = mc.mcdetector.Detectors(
        outer=mc.mcdetector.FiZ(Axis(-np.pi, np.pi, 16), Axis(-4e-3, 4e-3, 20)),
        specular=mc.mcdetector.Total())
    return mc.Mc(layers, mc.mcsource.Line((-8e-3, 0.3e-3, 0.0), (1.0, 0.05, 0.1)), det,
                 rnginit=1357, **kw), dict(rmax=20e-3)


ALL_CASES['mcvox_aniso_gauss_fluence'] = mcvox_aniso_gauss_fluence
GEOMETRY['mcvox_aniso_gauss_fluence'] = 'mcvox'
GOLDEN_RUN['mcvox_aniso_gauss_fluence'] = (2000, 16)
# the reference's mccyl Layers.check() rejects its own AnisotropicLayer
# (mccyl/mclayer/layer.py:938-941), so no golden vector can exist: CUDA vs oracle only
UNPINNED_CASES['mccyl_aniso_line_fiz'] = mccyl_aniso_line_fiz
UNPINNED_GEOMETRY['mccyl_aniso_line_fiz'] = 'mccyl'
UNPINNED_RUN['mccyl_aniso_line_fiz'] = (1500, 16)


# ---------------------------------------------------------------------------
# binary64 kernels (McDataTypesDouble, mcbase/mctypes.py:647-748,991-1044).  No C
# restatement exists for this family: the golden vectors (the reference kernel rendered
# in double precision, libm) pin the CUDA path directly.
def _double(mc):
    return mc.mctypes.McDataTypesDouble


def mcml_double_mhg_gauss_cart_flurz(mc, **kw):
    return mcml_mhg_gauss_cart_flurz(mc, types=_double(mc), **kw)


def mcml_double_lut_iso_radialpl_trace(mc, **kw):
    return mcml_lut_iso_radialpl_trace(mc, types=_double(mc), **kw)


def mcvox_double_gauss_fluence(mc, **kw):
    return mcvox_gauss_fluence(mc, types=_double(mc), **kw)


def mccyl_double_hg_line_fiz(mc, **kw):
    return mccyl_hg_line_fiz(mc, types=_double(mc), **kw)


def mcml_double_user_plugins(mc, **kw):
    """User-written phase function, source and detector (OpenCL-C fragments) in binary64."""
    return mcml_user_plugins(mc, types=_double(mc), **kw)


def mcml_double_user_surface_window(mc, **kw):
    """User-written surface layouts on both sample surfaces in binary64."""
    return mcml_user_surface_window(mc, types=_double(mc), **kw)


DOUBLE_CASES = {
    'mcml_double_user_plugins': mcml_double_user_plugins,
    'mcml_double_user_surface_window': mcml_double_user_surface_window,
    'mcml_double_mhg_gauss_cart_flurz': mcml_double_mhg_gauss_cart_flurz,
    'mcml_double_lut_iso_radialpl_trace': mcml_double_lut_iso_radialpl_trace,
    'mcvox_double_gauss_fluence': mcvox_double_gauss_fluence,
    'mccyl_double_hg_line_fiz': mccyl_double_hg_line_fiz,
}
DOUBLE_GEOMETRY = {name: name.split('_')[0] for name in DOUBLE_CASES}
DOUBLE_RUN = {name: (1500, 16) for name in DOUBLE_CASES}
