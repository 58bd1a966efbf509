"""Developer experiment: kernel time vs refill threshold."""
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1]; n = int(float(sys.argv[2]))
vals = [int(v) for v in sys.argv[3].split(',')]
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
sim = benchcfg.CONFIGS[name](mc)
for r in vals:
    sim.refill_lanes = r
    sim.run(10000, download=False)
    best = 1e9
    for i in range(3):
        sim.run(n, download=False)
        rr = sim.run_report
        best = min(best, rr['kernel_ms'])
    print(name, 'refill', r, 'kernel %.2f ms -> %.3e packets/s, %.1f iter/packet, %.3e iter/s regs %d grid %d' % (
        best, n/best*1e3, rr['iterations']/n, rr['iterations']/best*1e3,
        rr['kernel_attributes']['num_regs'], rr['grid']), flush=True)
_, fl, det = sim.run(n)
print('det', det.top.raw.sum()/n if det is not None else None, 'flu', fl.raw.sum()/n if fl is not None else None)
