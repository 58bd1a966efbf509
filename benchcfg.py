"""Concrete synthetic definitions of the BASELINE.json configs,
written against the simulator-module API so they can be instantiated with
``pyxopto_b200`` (bench, smoke, tests) or with the reference (oracle/build_ref.py).

C2 optical properties are the reference's ``xopto.materials.skin.Skin3()``
defaults evaluated at 550 nm with ``create_mc_layers`` (materials/skin/model.py:
1032-1066, 1121-1146) in this container; the subcutis is cut at 10 mm as SURVEY
8d prescribes.  They are plain numbers here because ``xopto.materials`` is input
generation, outside the accelerated path.
"""
import numpy as np

RNGINIT = 0x2545F4914F6CDD1D

SKIN3_550NM = [
    # (d [m], n, mua [1/m], mus [1/m], g)
    (1e-4, 1.334683329053798, 2493.6873583265165, 15079.876661995619, 0.9),   # epidermis
    (2e-3, 1.3865225836400437, 1001.2238137598148, 10079.982905825484, 0.8),  # dermis
    (1e-2, 1.3865225836400437, 792.2425907825923, 5039.991452912742, 0.8),    # subcutis
]
MHG_BETA = 0.9


def _fiber(mc):
    if mc.__name__.startswith('xopto'):
        from xopto.mcml.mcutil import fiber as fiberutil
        return fiberutil.MultimodeFiber(200e-6, 220e-6, 1.462, 0.22)
    return mc.mcsource.MultimodeFiber(200e-6, 220e-6, 1.462, 0.22)


def c1_slab(mc, rnginit=123456789, **kw):
    """BASELINE configs[0]: single slab, Line source, Radial reflectance."""
    Axis = mc.mcdetector.Axis
    pf = mc.mcpf.Hg(0.8)
    L = mc.mclayer.Layer
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0)),
        L(d=10e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0))])
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0.0, 10e-3, 1000)))
    sim = mc.Mc(layers, mc.mcsource.Line(), det, rnginit=rnginit, **kw)
    return sim


def c2_skin(mc, rnginit=RNGINIT, pf='mhg', **kw):
    """BASELINE configs[1]: 5-entry skin stack, UniformFiber source,
    SixAroundOne detector, MHg phase function, FluenceRz 250 x 500."""
    Axis = mc.mcdetector.Axis
    L = mc.mclayer.Layer

    def make_pf(g):
        return mc.mcpf.MHg(g, MHG_BETA) if pf == 'mhg' else mc.mcpf.Hg(g)

    stack = [L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=make_pf(0.0))]
    for d, n, mua, mus, g in SKIN3_550NM:
        stack.append(L(d=d, n=n, mua=mua, mus=mus, pf=make_pf(g)))
    stack.append(L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=make_pf(0.0)))
    fib = _fiber(mc)
    det = mc.mcdetector.Detectors(top=mc.mcdetector.SixAroundOne(fib, spacing=220e-6))
    flu = mc.mcfluence.FluenceRz(Axis(0.0, 5e-3, 250), Axis(0.0, 5e-3, 500))
    sim = mc.Mc(mc.mclayer.Layers(stack), mc.mcsource.UniformFiber(fib), det,
                fluence=flu, rnginit=rnginit, **kw)
    return sim


# algorithmic thread-operations per loop iteration: the
# per-unit figure of the roofline.  (ALU/FMA ops, MUFU ops)
OPS_PER_ITERATION = {
    'c1_slab': (85, 11),
    'c2_skin': (105, 12),
    'c3_vox': (83, 5.5),
}

CONFIGS = {'c1_slab': c1_slab, 'c2_skin': c2_skin}
GEOMETRY = {'c1_slab': 'mcml', 'c2_skin': 'mcml'}
PACKETS = {'c1_slab': 10**6, 'c2_skin': 10**9}


def c3_vox(mc, rnginit=RNGINIT, n=201, **kw):
    """BASELINE configs[2]: 201^3 voxel 2-layer skin with an embedded blood
    vessel (5 um voxels), GaussianBeam(sigma=50 um), Fluence deposition grid
    (parameters as in xopto/dataset/render/mcvox.py:39-46)."""
    A = mc.mcgeometry.Axis
    vs = 5e-6
    half = n/2*vs
    vox = mc.mcgeometry.Voxels(A(-half, half, n), A(-half, half, n), A(0.0, n*vs, n))
    M = mc.mcmaterial.Material
    nref = 1.337
    mats = mc.mcmaterial.Materials([
        M(n=nref, mua=0.0001e2, mus=1.0e2, pf=mc.mcpf.Hg(1.0)),
        M(n=nref, mua=16.5724e2, mus=375.9398e2, pf=mc.mcpf.Hg(0.9)),
        M(n=nref, mua=0.4585e2, mus=356.5406e2, pf=mc.mcpf.Hg(0.9)),
        M(n=nref, mua=230.5427e2, mus=93.9850e2, pf=mc.mcpf.Hg(0.9))])
    flu = mc.mcfluence.Fluence(vox.xaxis, vox.yaxis, vox.zaxis, mode='deposition')
    sim = mc.Mc(vox, mats, mc.mcsource.GaussianBeam(50e-6), fluence=flu,
                rnginit=rnginit, **kw)
    sim.rmax = 25e-3
    z, y, x = sim.voxels.meshgrid()
    m = sim.voxels.material
    m[z <= 100e-6] = 1
    m[z > 100e-6] = 2
    m[(x**2 + (z - 500e-6)**2) <= (100e-6)**2] = 3
    return sim


CONFIGS['c3_vox'] = c3_vox
GEOMETRY['c3_vox'] = 'mcvox'
PACKETS['c3_vox'] = 10**8


def c5_cyl(mc, rnginit=RNGINIT, **kw):
    """BASELINE configs[4], mccyl variant: one cylinder of radius
    5 mm (water, n = 1.337) in air, Line source along +x through the axis,
    FiZ(64 x 100) detector on the outer surface."""
    Axis = mc.mcdetector.Axis
    L = mc.mclayer.Layer
    pf = mc.mcpf.Hg(0.8)
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=pf),
        L(d=10e-3, n=1.337, mua=1e2, mus=100e2, pf=pf)])
    det = mc.mcdetector.Detectors(
        outer=mc.mcdetector.FiZ(Axis(-np.pi, np.pi, 64), Axis(-5e-3, 5e-3, 100)))
    sim = mc.Mc(layers, mc.mcsource.Line((-10e-3, 0.0, 0.0), (1.0, 0.0, 0.0)), det,
                rnginit=rnginit, **kw)
    sim.rmax = 25e-3
    return sim


CONFIGS['c5_cyl'] = c5_cyl
GEOMETRY['c5_cyl'] = 'mccyl'
PACKETS['c5_cyl'] = 10**7
# mcml AW + Hg (85, 11) with the two plane tests (4 ALU) replaced by the ray /
# cylinder quadratic of mccyl.template.c:147-209: a, b, c (9), 1/2a (1 + 1 MUFU),
# outer discriminant + root (6 + 1 MUFU), inner discriminant test (4)
OPS_PER_ITERATION['c5_cyl'] = (101, 13)


def c4_trace(mc, rnginit=RNGINIT, maxlen=512, **kw):
    """BASELINE configs[3] on the C1 slab: path-length resolved
    RadialPl reflectance (300 bins of 1 mm optical path ~ 3.3 ps), full Trace of
    every packet (maxlen 512, optical path length on) with a terminal-event
    filter selecting packets that leave through the top surface 0.5-1.5 mm from
    the source within 12.7 deg of the normal (a detector fibre), followed by
    ``sampling_volume`` on a 200^3 grid of 20 um voxels."""
    Axis = mc.mcdetector.Axis
    pf = mc.mcpf.Hg(0.8)
    L = mc.mclayer.Layer
    layers = mc.mclayer.Layers([
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0)),
        L(d=10e-3, n=1.33, mua=1e2, mus=100e2, pf=pf),
        L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0))])
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.RadialPl(Axis(0.0, 10e-3, 100), plaxis=Axis(0.0, 0.3, 300)))
    flt = mc.mctrace.Filter(z=(-float('inf'), 0.0), r=(0.5e-3, 1.5e-3, (0.0, 0.0)),
                            pz=(-1.0, -float(np.cos(np.deg2rad(12.7)))))
    tr = mc.mctrace.Trace(maxlen=maxlen, options=mc.mctrace.Trace.TRACE_ALL, plon=True,
                          filter=flt)
    sim = mc.Mc(layers, mc.mcsource.Line(), det, trace=tr, rnginit=rnginit, **kw)
    sim.rmax = 25e-3
    return sim


def c4_sampling_volume(mc):
    A = mc.mcsv.Axis
    return mc.mcsv.SamplingVolume(A(-2e-3, 2e-3, 200), A(-2e-3, 2e-3, 200), A(0.0, 4e-3, 200))


CONFIGS['c4_trace'] = c4_trace
GEOMETRY['c4_trace'] = 'mcml'
PACKETS['c4_trace'] = 10**6
# C1's loop (85, 11) plus the optical path length (1 FMA) and the trace event
# (address 3 IMAD + 2 STG.128); the binding resource is the 32 B trace record
# written per iteration (HBM), see bench.py
OPS_PER_ITERATION['c4_trace'] = (91, 11)
TRACE_BYTES_PER_ITERATION = 32


def c4_trace_vox(mc, rnginit=RNGINIT, maxlen=512, **kw):
    """BASELINE configs[3] on the voxel geometry ("time-resolved mcml/mcvox"): the C3
    medium (201^3 voxels of 5 um, 2-layer skin + vessel), Line source, path-length
    resolved RadialPl reflectance, full Trace of every packet (maxlen 512, one event
    per loop trip = voxel crossing or interaction, like the reference) with a
    terminal-event filter (top surface, 100-400 um from the source, within 30 deg of
    the normal), then ``sampling_volume`` on a 200^3 grid over the voxel box."""
    sim = c3_vox(mc, rnginit=rnginit, **kw)
    Axis = mc.mcdetector.Axis
    det = mc.mcdetector.Detectors(
        top=mc.mcdetector.RadialPl(Axis(0.0, 0.5e-3, 100), plaxis=Axis(0.0, 0.03, 300)))
    flt = mc.mctrace.Filter(z=(-float('inf'), 0.0), r=(100e-6, 400e-6, (0.0, 0.0)),
                            pz=(-1.0, -float(np.cos(np.deg2rad(30.0)))))
    tr = mc.mctrace.Trace(maxlen=maxlen, options=mc.mctrace.Trace.TRACE_ALL, plon=True,
                          filter=flt)
    vox, mats = sim.voxels, sim.materials
    material = np.array(sim.voxels.material, copy=True)
    sim2 = mc.Mc(vox, mats, mc.mcsource.Line(), det, trace=tr, rnginit=rnginit, **kw)
    sim2.rmax = 25e-3
    sim2.voxels.material[:] = material
    return sim2


def c4_sampling_volume_vox(mc):
    A = mc.mcsv.Axis
    half = 201/2*5e-6
    return mc.mcsv.SamplingVolume(A(-half, half, 200), A(-half, half, 200), A(0.0, 2*half, 200))


CONFIGS['c4_trace_vox'] = c4_trace_vox
GEOMETRY['c4_trace_vox'] = 'mcvox'
PACKETS['c4_trace_vox'] = 2*10**5
OPS_PER_ITERATION['c4_trace_vox'] = (83, 5.5)
SAMPLING_VOLUMES = {'c4_trace': c4_sampling_volume, 'c4_trace_vox': c4_sampling_volume_vox}


def c5_slab(mc, rnginit=RNGINIT, mua=1e2, musr=20e2, **kw):
    """BASELINE configs[4], mcml variant: semi-infinite water
    (n = 1.337) under air, Line source, Radial(0..5 mm, 500 bins) reflectance,
    rmax = 25 mm; one point of the 64 x 64 (mua, musr) grid, g = 0.8."""
    Axis = mc.mcdetector.Axis
    g = 0.8
    L = mc.mclayer.Layer
    layers = mc.mclayer.Layers([
        L(d=float('inf'), n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0)),
        L(d=float('inf'), n=1.337, mua=mua, mus=musr/(1.0 - g), pf=mc.mcpf.Hg(g)),
        L(d=float('inf'), n=1.0, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(0.0))])
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(Axis(0.0, 5e-3, 500)))
    sim = mc.Mc(layers, mc.mcsource.Line(), det, rnginit=rnginit, **kw)
    sim.rmax = 25e-3
    return sim


def c5_grid(n_mua=64, n_musr=64, g=0.8):
    """The (mua, musr) sweep of config 5 as Sweep descriptors."""
    return [{1: {'mua': float(mua), 'mus': float(musr/(1.0 - g))}}
            for mua in np.linspace(0.0, 5e2, n_mua) for musr in np.linspace(5e2, 35e2, n_musr)]


CONFIGS['c5_slab'] = c5_slab
GEOMETRY['c5_slab'] = 'mcml'
PACKETS['c5_slab'] = 10**7
OPS_PER_ITERATION['c5_slab'] = (85, 11)


# ---- the reference's own acceptance / performance workload ----------------------
# xopto/mcml/test/validate.py:348-445 (SingleLayerUniformFiberRadial) is the workload
# behind every number printed by test/performance.py:116-162: one 8 mm layer (n 1.33,
# Hg g 0.85) under a fiber-glass half-space (n 1.452), UniformFiberNI(200 um, NA 0.22),
# Radial(RadialAxis(0, 1.7 mm, 340), cosmin = cos(asin(NA/ncore))), rmax 5 mm, and a
# 20 x 20 grid mua in linspace(1, 2500, 20) 1/m x musr in linspace(500, 6000, 20) 1/m
# with 1e7 packets each.  Parameters are the ones stored in the reference's
# test/reference/cuda_singlelayer_uniformfiber.pkl (tests/golden/validate_vectors.npz).
VALIDATE_G = 0.85
VALIDATE_NCORE = 1.452
VALIDATE_NA = 0.22


def validate_uniformfiber(mc, rnginit=RNGINIT, mua=1.0, musr=500.0, **kw):
    L = mc.mclayer.Layer
    g = VALIDATE_G
    layers = mc.mclayer.Layers([
        L(d=float('inf'), n=VALIDATE_NCORE, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(g)),
        L(d=8e-3, n=1.33, mua=mua, mus=musr/(1.0 - g), pf=mc.mcpf.Hg(g)),
        L(d=float('inf'), n=1.33, mua=0.0, mus=0.0, pf=mc.mcpf.Hg(g))])
    if mc.__name__.startswith('xopto'):
        from xopto.mcml.mcutil import fiber as fiberutil
        fib = fiberutil.MultimodeFiber(200e-6, 200e-6, VALIDATE_NCORE, VALIDATE_NA)
    else:
        fib = mc.mcsource.MultimodeFiber(200e-6, 200e-6, VALIDATE_NCORE, VALIDATE_NA)
    cosmin = float(np.cos(np.arcsin(VALIDATE_NA/VALIDATE_NCORE)))
    det = mc.mcdetector.Detectors(top=mc.mcdetector.Radial(
        mc.mcdetector.RadialAxis(0.0, 1.7e-3, 340), cosmin=cosmin))
    sim = mc.Mc(layers, mc.mcsource.UniformFiberNI(fib), det, rnginit=rnginit, **kw)
    sim.rmax = 5e-3
    return sim


def validate_grid(n_mua=20, n_musr=20, g=VALIDATE_G, layer=1):
    """The 400 (mua, musr) points of the acceptance suite as Sweep descriptors
    (mua-major, musr fastest: np.meshgrid(..., indexing='ij') of validate.py:413)."""
    return [{layer: {'mua': float(mua), 'mus': float(musr/(1.0 - g))}}
            for mua in np.linspace(1.0, 2500.0, n_mua)
            for musr in np.linspace(500.0, 6000.0, n_musr)]


CONFIGS['validate_uniformfiber'] = validate_uniformfiber
GEOMETRY['validate_uniformfiber'] = 'mcml'
PACKETS['validate_uniformfiber'] = 10**7
OPS_PER_ITERATION['validate_uniformfiber'] = (85, 11)
# published by the reference for exactly this workload (test/performance.py:134-135):
# 400 x 1e7 packets in 5.7 s on an NVIDIA RTX A6000 (OpenCL)
VALIDATE_PUBLISHED_SECONDS = 5.7
VALIDATE_PUBLISHED_PACKETS_PER_S = 400*1e7/5.7
