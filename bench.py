#!/usr/bin/env python
"""Headline benchmark: photon packets/s of the mcml 5-layer skin configuration
(BASELINE.json configs[1], concrete definition in benchcfg.c2_skin / SURVEY 8d).

    python bench.py --gpus N --steps K --warmup W [--config c2_skin] [--packets P]
    python bench.py --impl reference ...      (the reference's own kernel on host cores)

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE).  A step simulates
P packets per GPU (default 1.25e8 = 1e9 / 8, weak scaling) with a rank-specific
MWC seed set and, for N > 1, combines the 64-bit fixed-point accumulators with a
single NCCL all-reduce.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import benchcfg  # noqa: E402

SM_COUNT_B200 = 148
FP32_LANES_PER_SM = 128
SFU_LANES_PER_SM = 16


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'sm_max_mhz': 1965.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampler running during the timed region."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)),
                'power_w': float(np.median(power)), 'samples': len(sm),
                'reasons': sorted(reasons)}


def cpu_reference_run(config: str, sample_packets: int, threads: int, variant: str = None):
    """The reference's own kernel (oracle/_ref, built here from the rendered
    reference text, on inputs packed by the reference's host layer) or, when that
    build is absent, the oracle port, on host cores with the dynamic schedule.
    Returns (packets/s, kind, seconds, compiler flags).  Imports nothing of
    pyxopto_b200 when the reference build is present."""
    oracle_dir = os.path.join(ROOT, 'oracle')
    if oracle_dir not in sys.path:
        sys.path.insert(0, oracle_dir)
    import refbench
    return refbench.run(config, benchcfg.GEOMETRY[config], sample_packets, threads, variant)


def cpu_reference_variant(config: str, pilot: int, threads: int):
    """The faster usable build of the reference kernel (IEEE or -ffast-math)."""
    oracle_dir = os.path.join(ROOT, 'oracle')
    if oracle_dir not in sys.path:
        sys.path.insert(0, oracle_dir)
    import refbench
    return refbench.fastest_variant(config, benchcfg.GEOMETRY[config], pilot, threads)[0]


METRIC_NAMES = {'c2_skin': 'mcml 5-layer skin', 'c3_vox': 'mcvox 201^3 fluence',
                'validate_uniformfiber': "reference's validate.SingleLayerUniformFiberRadial "
                                         'performance suite, 400 x 1e7 packets'}
WORKLOADS = {
    'c2_skin': 'mcml 5-layer skin (Skin3 @550nm), UniformFiber + SixAroundOne, '
               'MHg(beta=0.9), FluenceRz 250x500',
    'c1_slab': 'mcml single slab mua=1/cm mus=100/cm g=0.8 n=1.33, Line + Radial',
    'c3_vox': 'mcvox 201^3 voxel 2-layer skin + blood vessel (5 um voxels), '
              'GaussianBeam sigma 50 um, Fluence deposition grid',
    'c4_trace': 'mcml slab, RadialPl 100x300 (path-length resolved) + Trace maxlen 512 of every '
                'packet + device filter + sampling_volume 200^3 (e2e: the accepted rows feed '
                'sampling_volume on the device, one SamplingVolume accumulates over the steps '
                'on the device; the host receives detectors and counts per step and the '
                'grid once, inside the timed region)',
    'c4_trace_vox': 'mcvox 201^3 skin + vessel, Line source, Trace maxlen 512 of every packet + '
                    'device filter + sampling_volume 200^3',
    'c5_slab': 'mcml semi-infinite n=1.337 under air, (mua, musr) sweep point(s), g=0.8, '
               'Line + Radial 500, rmax 25 mm',
    'c5_cyl': 'mccyl single cylinder r=5 mm n=1.337 mua=1/cm mus=100/cm g=0.8, '
              'Line + FiZ 64x100',
    'validate_uniformfiber': 'mcml 8 mm layer n=1.33 Hg(0.85) under n=1.452, UniformFiberNI(200 um, '
                             'NA 0.22) + Radial(RadialAxis 340 x 5 um, cosmin), rmax 5 mm; sweep of '
                             '20 x 20 (mua, musr) points x 1e7 packets '
                             '(xopto/mcml/test/validate.py:348-445, test/performance.py)',
}
DEFAULT_PACKETS = {'c2_skin': 1.25e8, 'c3_vox': 1e8, 'c4_trace': 1e6, 'c4_trace_vox': 2e5}
TRACE_CONFIGS = ('c4_trace', 'c4_trace_vox')


def atomic_peaks():
    """Measured accumulator-path rates (tools/atomic_probe.cu run on a B200, output
    committed as profiles/atomic_probe_r02.json): deposits/s of the shared-memory
    lo/hi window (ATOMS) and of RED.E.ADD.64 into a 201^3 grid in L2 (uniform)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'atomic_probe_r02.json')) as f:
            res = json.load(f)['results']
        atoms = max(r['G_deposits_per_s'] for r in res
                    if r['path'] == 'atoms_lohi' and r['bins'] >= 4096 and not r['peaked'])
        red = max(r['G_deposits_per_s'] for r in res
                  if r['path'] == 'redg64' and r['bins'] == 8120601 and not r['peaked'])
        return atoms*1e9, red*1e9, 'profiles/atomic_probe_r02.json'
    except (OSError, ValueError, KeyError):
        return 1.2e12, 1.8e11, 'constants (probe output not found)'


def reference_arm(config, args, ncores):
    sample = int(args.cpu_sample or {'c2_skin': 3e5, 'c3_vox': 2e4, 'c4_trace': 2e4,
                                     'c4_trace_vox': 1e4}.get(config, 5e4))
    times = []
    kind, flags = 'port', ''
    variant = cpu_reference_variant(config, max(sample//4, 1000), ncores)
    for i in range(args.warmup + args.steps):
        pps, kind, secs, flags = cpu_reference_run(config, sample, ncores, variant)
        if i >= args.warmup:
            times.append(secs)
    total = sum(times)
    value = sample*len(times)/total
    return {
        'impl': 'reference',
        'metric': 'photon packets/s ({})'.format(METRIC_NAMES.get(config, config)),
        'value': value, 'unit': 'packets/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3*total/len(times), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS.get(config, config), 'packets_per_step': sample,
                   'schedule': 'dynamic atomic packet counter, one work-item per host thread',
                   'inputs': 'structs / LUTs / seeds packed by the reference host layer '
                             '(oracle/_ref/inputs_{}.npz)'.format(config)},
        'cpu_baseline': {'value': value, 'unit': 'packets/s', 'cores': ncores,
                         'kind': kind, 'flags': flags,
                         'sample': '{} packets per step, {} timed steps'.format(sample, len(times))},
        'e2e': {'value': value, 'unit': 'packets/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


def native_arm(config, args, env, cpu_baseline_wanted=True):
    """One bench line (dict, rank 0; None elsewhere) of ``config``."""
    import importlib
    from pyxopto_b200.cu import abi
    from pyxopto_b200 import parallel
    rank, world, local_rank = env['rank'], env['world'], env['local_rank']
    torch, dist = env.get('torch'), env.get('dist')
    ncores = os.cpu_count() or 1
    geom = benchcfg.GEOMETRY[config]
    mc = importlib.import_module('pyxopto_b200.{}.mc'.format(geom))
    packets = int(args.packets or DEFAULT_PACKETS.get(config, 1e7))
    sim = benchcfg.CONFIGS[config](mc, rnginit=parallel.seed_for_rank(benchcfg.RNGINIT, rank),
                                   cl_devices=local_rank)
    n_sweep = int(args.sweep)
    if config == 'validate_uniformfiber' and not n_sweep:
        n_sweep = 400
    reducer = None
    if world > 1 and not n_sweep:
        # packets sharded over the ranks: one all-reduce of the accumulators per step,
        # stream-ordered on the engine's stream
        reducer = parallel.NcclAccumulatorReducer(local_rank)
        sim._reduce_hook = reducer

    def barrier():
        sim._stream.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    # one step of the configuration, device-resident (`value`) and through the
    # public API with host results (`e2e`)
    sv_ms = []
    launches_per_step = 1
    if config in TRACE_CONFIGS:
        launches_per_step = 5

        def step_device():
            sim.run(packets, download=False)
            rr = dict(sim.run_report)
            sim.filter_trace_on_device(packets, download=False)
            sim.sampling_volume(None, benchcfg.SAMPLING_VOLUMES[config](mc), download=False)
            sv_ms.append(sim.run_report['sv_kernel_ms'] + sim.run_report['filter_ms'])
            return rr

        # the public workflow of a sampling-volume study: every batch goes through
        # run() (detectors and counts to the host) and sampling_volume(trace, sv) into
        # ONE SamplingVolume object; its grid accumulates on the device (exact 64-bit
        # sums) and reaches the host once, when `sv.data` is read after the last batch
        # (inside the timed region, see `finish_e2e`)
        sim.lazy_sampling_volume = True
        sv_acc = [benchcfg.SAMPLING_VOLUMES[config](mc)]

        trace_d2h = {}

        def step_e2e():
            trace, fluence, detectors = sim.run(packets)
            trace_d2h['detectors'] = int(sim.cl_rw_accumulator_allocator.size)*8
            trace_d2h['counts'] = 4*int(sim.run_report.get('filter_accepted', 0)) + 8
            sim.sampling_volume(trace, sv_acc[0])
            trace_d2h['grid'] = int(sim.cl_rw_accumulator_allocator.size)*8
            return detectors, fluence, 0.0

        def finish_e2e():
            total = float(sv_acc[0].data.sum())
            sv_acc[0] = benchcfg.SAMPLING_VOLUMES[config](mc)
            return total
    elif n_sweep:
        # a pipelined sweep over (mua, musr): configurations dealt by a fixed pseudo-random permutation to the
        # ranks, no collective on the data path, one gather of the rows in e2e
        from pyxopto_b200 import mcsweep
        grid_cfgs = benchcfg.validate_grid() if config == 'validate_uniformfiber' \
            else benchcfg.c5_grid()
        total_cfgs = n_sweep*world
        if total_cfgs >= len(grid_cfgs):
            reps = (total_cfgs + len(grid_cfgs) - 1)//len(grid_cfgs)
            sweep_cfgs = (grid_cfgs*reps)[:total_cfgs]
        else:
            stride = max(len(grid_cfgs)//total_cfgs, 1)
            sweep_cfgs = grid_cfgs[::stride][:total_cfgs]
        sweep = mcsweep.Sweep(sim, rank, world)
        # N > 1: a short pilot (2000 packets per configuration, the ranks pilot disjoint
        # slices, one all-reduce of the loop-trip counts; outside the timed region, like
        # the kernel build) lets the configurations be dealt longest-first
        sweep_costs = sweep.pilot_costs(sweep_cfgs) if world > 1 else None
        launches_per_step = n_sweep
        ev_a, ev_b = abi.Event(sim.cl_context), abi.Event(sim.cl_context)

        def step_device():
            ev_a.record(sim._stream)
            sweep.run(sweep_cfgs, packets, costs=sweep_costs)
            ev_b.record(sim._stream)
            sim._stream.synchronize()
            rr = dict(sim.run_report)
            rr['kernel_ms'] = ev_a.elapsed_ms(ev_b)       # all kernels of the sweep
            rr['iterations'] = int(sweep.report['iterations'].sum())
            rr.setdefault('kernel_attributes', {'num_regs': None})
            return rr

        def step_e2e():
            idx, rows = sweep.run(sweep_cfgs, packets, costs=sweep_costs)
            if world > 1:
                rows = sweep.gather(idx, rows, len(sweep_cfgs))
            refl = sweep.detector(rows, sim.detectors.top if geom != 'mccyl'
                                  else sim.detectors.outer, packets)
            return None, None, float(np.sum(refl))
    else:
        def step_device():
            sim.run(packets, download=False)
            return sim.run_report

        def step_e2e():
            if world > 1:
                # the root rank receives the reduced results; the others only simulate
                trace, fluence, detectors = parallel.run_sharded(
                    sim, packets*world, rank, world, reducer=reducer, root=0)
            else:
                trace, fluence, detectors = sim.run(packets)
            return detectors, fluence, 0.0

    finish = locals().get('finish_e2e')
    # warm-up (also builds/loads the kernel)
    for _ in range(max(args.warmup, 1)):
        step_device()
    # (two passes, the first result still held while the second is produced - as in the
    # timed loop below: large grids come back in page-locked buffers of a small pool that
    # the results own, and the second 65 MB buffer of C3 is allocated here, not inside
    # the timed region)
    held = step_e2e()
    held = step_e2e()
    del held
    if finish is not None:
        finish()
    barrier()

    # ---- device-resident loop: `value` -------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_ms, iters = [], []
    ev_start, ev_stop = abi.Event(sim.cl_context), abi.Event(sim.cl_context)
    barrier()
    t0 = time.perf_counter()
    ev_start.record(sim._stream)
    for _ in range(args.steps):
        rr_step = step_device()
        kernel_ms.append(rr_step['kernel_ms'])
        iters.append(rr_step['iterations'])
    ev_stop.record(sim._stream)
    barrier()
    t1 = time.perf_counter()
    loop_ms_events = ev_start.elapsed_ms(ev_stop)
    clocks = sampler.stop() if rank == 0 else None
    loop_s = t1 - t0

    # ---- end-to-end loop through the public API: `e2e` ----------------------------
    barrier()
    t2 = time.perf_counter()
    checksum = touched = 0.0
    for _ in range(args.steps):
        detectors, fluence, extra = step_e2e()
        # every step's results are host arrays when step_e2e returns (the download is
        # inside Mc.run); the loop reads the detector bins and one value per 4 KB page of
        # the grid - summing all 8.1e6 float64 cells of C3 with NumPy took 4 ms of the
        # 80 ms step and is not part of the path.  The full checksum of the last step's
        # result follows the timed region.
        touched += sum(float(d.raw.sum()) for d in (detectors or ()) if hasattr(d, 'raw')) + \
            (float(fluence.raw.reshape(-1)[::512].sum()) if fluence is not None else 0.0) + extra
    if finish is not None:
        checksum += finish()
    barrier()
    t3 = time.perf_counter()
    e2e_s = t3 - t2
    checksum += sum(float(d.raw.sum()) for d in (detectors or ()) if hasattr(d, 'raw')) + \
        (float(fluence.raw.sum()) if fluence is not None else 0.0) + extra
    assert touched == touched          # (not NaN: the sampled reads really happened)
    from pyxopto_b200.cl import cltypes
    P = sim._packed
    h2d = sum(len(cltypes.raw_bytes(P[k])) for k in P if P[k] is not None) + 16
    d2h = int(sim.cl_rw_accumulator_allocator.size)*8 + 16
    if n_sweep:
        h2d, d2h = h2d*n_sweep, d2h*n_sweep
    if config in TRACE_CONFIGS:
        # per step: detector bins, counts of the accepted packets (their rows stay on the
        # device for sampling_volume), the total weight; the float64 grid once per run
        d2h = trace_d2h['detectors'] + trace_d2h['counts'] + 8 + trace_d2h['grid']//args.steps

    if world > 1:
        t = torch.tensor([loop_s, e2e_s, loop_ms_events], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loop_s, e2e_s, loop_ms_events = [float(v) for v in t.tolist()]

    per_step = packets*max(n_sweep, 1)
    total_packets = per_step*world*args.steps
    value = total_packets/loop_s
    e2e_value = total_packets/e2e_s
    if rank != 0:
        return None

    peaks, peaks_src = measured_peaks()
    sm_mhz_max = float(peaks.get('sm_max_mhz', 1965.0))
    info = sim.cl_context.info
    sms = info['multiprocessor_count']
    issue_peak = sms*FP32_LANES_PER_SM*sm_mhz_max*1e6     # thread-instr/s
    sfu_peak = sms*SFU_LANES_PER_SM*sm_mhz_max*1e6
    alu_ops, sfu_ops = benchcfg.OPS_PER_ITERATION[config]
    k_ms = float(np.mean(kernel_ms))
    iter_per_launch = float(np.mean(iters))
    achieved = iter_per_launch*(alu_ops + sfu_ops)/(k_ms*1e-3)
    achieved_sfu = iter_per_launch*sfu_ops/(k_ms*1e-3)
    # DRAM bytes of the kernel from the committed ncu capture (per launch, at the
    # capture's launch size; the working set of this path lives in L2)
    traffic, traffic_note, tr = None, None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            tr = json.load(f).get(config)
        if tr:
            traffic = tr['dram_bytes']
            traffic_note = 'dram read+write bytes of one McKernel launch of {:.0e} packets ' \
                           '({})'.format(tr['launch_packets'], tr['capture'])
    except (OSError, ValueError, KeyError):
        pass
    atoms_peak, red_peak, atomic_src = atomic_peaks()
    dep_rate = iter_per_launch/(k_ms*1e-3)
    # Where the engine does the physics of one reference loop trip with fewer instructions
    # than the reference's loop (mcvox: one ray per flight, clearance shortcut), the
    # reference-unit figure above would exceed the peak.  The fraction reported is then
    # the EXECUTED thread-instructions per trip (committed ncu capture) over the lane
    # slots of the SMs - issue utilisation x active lanes / 32 - and the reference-unit
    # figure is kept beside it.
    reference_units = None
    executed = (tr or {}).get('executed_thread_instr_per_iteration')
    if executed:
        reference_units = {'achieved': achieved/1e9, 'frac': achieved/issue_peak,
                           'ops_per_iteration': alu_ops + sfu_ops,
                           'note': 'thread-ops the reference loop needs for the same trips'}
        achieved = iter_per_launch*float(executed)/(k_ms*1e-3)
    roofline = {
        'bound': 'issue', 'achieved': achieved/1e9, 'peak': issue_peak/1e9,
        'unit': 'G thread-instr/s', 'frac': achieved/issue_peak,
        'traffic': traffic, 'traffic_note': traffic_note,
        'kernel': 'McKernel', 'kernel_ms': k_ms,
        'iterations_per_launch': iter_per_launch,
        'iterations_per_packet': iter_per_launch/per_step,
        'algorithmic_ops_per_iteration': {'alu_fma': alu_ops, 'mufu': sfu_ops},
        'executed_thread_instr_per_iteration': executed,
        'reference_units': reference_units,
        'sfu': {'achieved': achieved_sfu/1e9, 'peak': sfu_peak/1e9,
                'frac': achieved_sfu/sfu_peak},
        # accumulator ("atomic") roofline, SURVEY 8d: deposits per second against
        # the rates measured with tools/atomic_probe.cu on B200; deposits per
        # iteration = 1 (C2, AW) / 0.15 (C3, oracle statistics)
        'atomic': ({'achieved': dep_rate, 'peak': atoms_peak, 'frac': dep_rate/atoms_peak,
                    'unit': 'deposits/s', 'peak_source': atomic_src,
                    'path': 'shared-memory window (ATOMS lo/hi), 2.5 % RED.E.ADD.64'}
                   if config == 'c2_skin' else
                   {'achieved': 0.15*dep_rate, 'peak': red_peak, 'frac': 0.15*dep_rate/red_peak,
                    'unit': 'deposits/s', 'peak_source': atomic_src,
                    'path': 'RED.E.ADD.64 to L2'}
                   if config == 'c3_vox' else None),
        'peak_source': 'sm_max_mhz of MEASURED_PEAKS.json ({}) x {} SMs x {} FP32 lanes; '
                       'hbm is not the bound of this path (working set < 2 MB / L2-resident)'.format(
                           peaks_src, sms, FP32_LANES_PER_SM),
        'kernel_share_of_step': k_ms*args.steps/(loop_ms_events if loop_ms_events > 0 else 1),
    }
    if config in TRACE_CONFIGS:
        # the trace stream binds this configuration: one 32 B event record per
        # loop iteration, written once
        hbm_peak = float(peaks.get('hbm_gbs', peaks.get('hbm_gbps', 6650.0)))
        bytes_per_launch = iter_per_launch*benchcfg.TRACE_BYTES_PER_ITERATION
        roofline.update({
            'bound': 'hbm', 'achieved': bytes_per_launch/(k_ms*1e-3)/1e9,
            'peak': hbm_peak, 'unit': 'GB/s',
            'frac': bytes_per_launch/(k_ms*1e-3)/1e9/hbm_peak,
            'algorithmic_bytes_per_iteration': benchcfg.TRACE_BYTES_PER_ITERATION,
            'issue': {'achieved': achieved/1e9, 'peak': issue_peak/1e9,
                      'frac': achieved/issue_peak},
            'filter_plus_sampling_volume_ms': float(np.mean(sv_ms)) if sv_ms else None,
            'peak_source': 'hbm_gbs of MEASURED_PEAKS.json ({})'.format(peaks_src),
        })
    cpu_baseline = None
    if cpu_baseline_wanted and not args.no_cpu_baseline and world == 1:
        try:
            variant = None
            if args.cpu_sample:
                sample = int(args.cpu_sample)
            else:
                # pilot run, then a sample sized for ~12 s of CPU work
                pilot = int(5e3 if config in TRACE_CONFIGS else 5e4)
                variant = cpu_reference_variant(config, pilot, ncores)
                pps0 = cpu_reference_run(config, pilot, ncores, variant)[0]
                sample = int(min(max(pps0*12.0, pilot), 1e5 if config in TRACE_CONFIGS else 2e8))
            pps, kind, secs, flags = cpu_reference_run(config, sample, ncores, variant)
            cpu_baseline = {'value': pps, 'unit': 'packets/s', 'cores': ncores,
                            'kind': kind, 'flags': flags,
                            'sample': '{} packets of the same workload in {:.1f} s'.format(
                                sample, secs)}
        except Exception as exc:    # the baseline must never take the bench down
            cpu_baseline = {'value': None, 'unit': 'packets/s', 'cores': ncores,
                            'kind': 'port', 'sample': 'failed: {!r}'.format(exc)}
    rr = sim.run_report
    vs_baseline = None
    if config == 'validate_uniformfiber' and n_sweep == 400 and packets == 10**7:
        # the one number the reference publishes for this path: 400 x 1e7 packets in
        # 5.7 s on an RTX A6000 (test/performance.py:134-135) - end to end
        vs_baseline = e2e_value/world/benchcfg.VALIDATE_PUBLISHED_PACKETS_PER_S
    line = {
        'metric': 'photon packets/s ({})'.format(METRIC_NAMES.get(config, config)),
        'value': value, 'unit': 'packets/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3*loop_s/args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': vs_baseline, 'dtype': 'f32', 'data': 'synthetic',
        'config': {
            'workload': WORKLOADS.get(config, config), 'packets_per_gpu_per_step': per_step,
            'global_packets_per_step': per_step*world,
            'sweep_configs_per_gpu_per_step': n_sweep or None,
            'parallelism': ('{} configurations per GPU dealt longest-first by pilot cost over {} GPU(s), no '
                            'collective on the data path'.format(n_sweep, world) if n_sweep else
                            'packets sharded over {} GPU(s), disjoint MWC seed sets{}'.format(
                                world, ', 1 stream-ordered NCCL all-reduce of the uint64 '
                                'accumulators per step' if world > 1 else '')),
            'mode': 'throughput (MUFU math, dynamic chunked packet counter)',
            'l2': 'inputs are < 2 MB of constants; every step re-zeroes and rewrites '
                  'the accumulator grid, no cached outputs are reused',
            'grid': rr['grid'], 'block': rr['block'],
            'registers': rr.get('kernel_attributes', {}).get('num_regs'),
        },
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'packets/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': 1e3*e2e_s/args.steps,
                'checksum': checksum,
                'results_on': 'rank 0 (NCCL-reduced accumulators)' if world > 1 else 'host'},
        'gpu_launches': args.steps*launches_per_step,
        'roofline': roofline,
        'cpu_baseline': cpu_baseline,
    }
    if config == 'validate_uniformfiber':
        line['suite_seconds'] = {'device': loop_s/args.steps, 'e2e': e2e_s/args.steps,
                                 'published_rtx_a6000': benchcfg.VALIDATE_PUBLISHED_SECONDS}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--config', default=None, choices=sorted(benchcfg.CONFIGS))
    ap.add_argument('--packets', type=float, default=None,
                    help='packets per GPU per step (default 1.25e8 for c2_skin)')
    ap.add_argument('--cpu-sample', type=float, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true',
                    help='default run only: skip the mcvox 201^3 block')
    ap.add_argument('--sweep', type=int, default=0,
                    help='c5_slab / c5_cyl / validate_uniformfiber: configurations per GPU '
                         'per step, run as a pipelined sweep')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    ncores = os.cpu_count() or 1
    # BASELINE.json's metric names two workloads: "mcml skin, mcvox 201^3 fluence".
    # Without --config the line is the mcml skin measurement and carries the mcvox
    # one, measured the same way in the same run, as `secondary`.
    default_run = args.config is None
    config = args.config or 'c2_skin'

    if args.impl == 'reference':
        if rank != 0:
            return 0
        line = reference_arm(config, args, ncores)
        if default_run and not args.no_secondary:
            line['secondary'] = reference_arm('c3_vox', args, ncores)
        print(json.dumps(line))
        return 0

    env = {'rank': rank, 'world': world, 'local_rank': local_rank}
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        env.update(torch=torch, dist=dist)
    line = native_arm(config, args, env)
    if default_run and not args.no_secondary:
        second = native_arm('c3_vox', args, env)
        if line is not None:
            line['secondary'] = second
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        env['dist'].destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
