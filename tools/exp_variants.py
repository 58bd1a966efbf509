"""Developer experiment: kernel time of C2 variants (fluence on/off, pf kinds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
from pyxopto_b200.mcml import mc

def build(fluence=True, pf='mhg', source='fiber', det=True):
    Axis = mc.mcdetector.Axis
    L = mc.mclayer.Layer
    def make_pf(g):
        return mc.mcpf.MHg(g, benchcfg.MHG_BETA) if pf == 'mhg' else mc.mcpf.Hg(g)
    stack = [L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=make_pf(0.0))]
    for d, n, mua, mus, g in benchcfg.SKIN3_550NM:
        stack.append(L(d=d, n=n, mua=mua, mus=mus, pf=make_pf(g)))
    stack.append(L(d=0.0, n=1.0, mua=0.0, mus=0.0, pf=make_pf(0.0)))
    fib = benchcfg._fiber(mc)
    d = mc.mcdetector.Detectors(top=mc.mcdetector.SixAroundOne(fib, spacing=220e-6)) if det else None
    flu = mc.mcfluence.FluenceRz(Axis(0.0, 5e-3, 250), Axis(0.0, 5e-3, 500)) if fluence else None
    src = mc.mcsource.UniformFiber(fib) if source == 'fiber' else mc.mcsource.Line()
    return mc.Mc(mc.mclayer.Layers(stack), src, d, fluence=flu, rnginit=benchcfg.RNGINIT)

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000000
for kw in (dict(), dict(fluence=False), dict(pf='hg'), dict(source='line'),
           dict(source='line', fluence=False), dict(source='line', fluence=False, det=False, pf='hg')):
    sim = build(**kw); sim.refill_lanes = 6 if kw.get("source") != "line" else 1
    sim.run(10000, download=False)
    best = 1e9
    for i in range(3):
        sim.run(n, download=False)
        rr = sim.run_report
        best = min(best, rr['kernel_ms'])
    print(kw, 'kernel %.2f ms -> %.3e packets/s, %.1f iter/packet, %.3e iter/s regs %d' % (
        best, n/best*1e3, rr['iterations']/n, rr['iterations']/best*1e3,
        rr['kernel_attributes']['num_regs']), flush=True)
