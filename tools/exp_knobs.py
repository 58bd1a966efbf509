"""Developer experiment: kernel time vs knobs. usage: exp_knobs.py config n 'k=v,k=v;k=v'"""
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
name = sys.argv[1]; n = int(float(sys.argv[2]))
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
for spec in sys.argv[3].split(';'):
    sim = benchcfg.CONFIGS[name](mc)
    kw = {}
    for kv in filter(None, spec.split(',')):
        k, v = kv.split('=')
        if k == 'wgsize': kw['wgsize'] = int(v)
        elif k.startswith('D'): sim._cl_build_options.append('-%s=%s' % (k, v))
        else: setattr(sim, k, float(v) if '.' in v else int(v))
    sim.run(10000, download=False, **kw)
    best = 1e9
    for i in range(3):
        sim.run(n, download=False, **kw)
        rr = sim.run_report
        best = min(best, rr['kernel_ms'])
    print(name, spec, 'kernel %.2f ms -> %.3e packets/s, %.3e iter/s regs %d grid %d block %d smem %d win %s' % (
        best, n/best*1e3, rr['iterations']/best*1e3, rr['kernel_attributes']['num_regs'], rr['grid'], rr['block'],
        rr['shared_bytes'], rr['fluence_window']), flush=True)
