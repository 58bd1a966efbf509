"""Developer timing probe of one bench configuration (also the process ncu profiles,
tools/gpu_ncu.sh): python tools/probe_config.py config n [wgsize]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchcfg
import importlib
name = sys.argv[1]; n = int(float(sys.argv[2]))
kw = {'wgsize': int(sys.argv[3])} if len(sys.argv) > 3 else {}
mc = importlib.import_module("pyxopto_b200.%s.mc" % benchcfg.GEOMETRY[name]); sim = benchcfg.CONFIGS[name](mc)
if os.environ.get('XO_REFILL'):
    sim.refill_lanes = int(os.environ['XO_REFILL'])
if os.environ.get('XO_MIN_BLOCKS'):
    sim.min_blocks = int(os.environ['XO_MIN_BLOCKS'])
if os.environ.get('XO_POOL_SLOTS'):
    sim.pool_slots = int(os.environ['XO_POOL_SLOTS'])
if any(k.startswith('XO_POOL_') and k != 'XO_POOL_SLOTS' for k in os.environ):
    sim.pool_tuning = {k: int(v) for k, v in os.environ.items()
                       if k.startswith('XO_POOL_') and k != 'XO_POOL_SLOTS'}
sim.run(10000, download=False, **kw)
for i in range(3):
    sim.run(n, download=False, **kw)
    rr = sim.run_report
    print(name, 'n=%.0e kernel %.2f ms -> %.3e packets/s, %.1f iter/packet, %.3e iter/s | grid %d x %d regs %d smem %d priv %d win %s' % (n, rr['kernel_ms'], n/rr['kernel_ms']*1e3, rr['iterations']/n, rr['iterations']/rr['kernel_ms']*1e3, rr['grid'], rr['block'], rr['kernel_attributes']['num_regs'], rr['shared_bytes'], rr['private_bins'], rr['fluence_window']), flush=True)
t=time.perf_counter(); tr, fl, det = sim.run(n, **kw); dt=time.perf_counter()-t
print('e2e run %.3f s -> %.3e packets/s' % (dt, n/dt), {k: (round(v,4) if isinstance(v,float) else v) for k,v in sim.run_report.items() if k in ('upload','execution','download','build')})
print("det", None if det is None else det.top.raw.sum()/n)
if fl is not None: print('fluence total', fl.raw.sum()/n)
