"""Generates tests/golden/*.npz with the REFERENCE itself (this container only).

For every case in tests/cases.py the reference package packs its structs and
renders its kernel; the kernel text is compiled unchanged for the CPU
(oracle/refkernel.py) and run under the static block schedule.  Stored per case:
the raw packed struct bytes, the float LUT pool, and the resulting accumulator /
int / float buffers + advanced RNG states.  These pin (a) the host mirror's
packing and (b) the oracle restatement on machines without /root/reference.

    python tests/golden/make_golden.py [case ...]
    python tests/golden/make_golden.py bench [config ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)

import ref_env  # noqa: E402


def make_lut():
    ref_env.activate()
    from xopto import pf
    params, lut = pf.Hg(0.8).mclut(2000)
    np.savez_compressed(os.path.join(HERE, 'lut_hg08_2000.npz'),
                        params=np.asarray(params, np.float64),
                        lut=np.asarray(lut, np.float64))


def make_bench(names):
    """tests/golden/bench_<config>.npz: the bench configurations (benchcfg.py) at
    their real size through the reference kernel, static block schedule."""
    ref_env.activate()
    import cases
    from refkernel import RefKernel
    import importlib
    for name in names or sorted(cases.BENCH_RUN):
        geom = cases.bench_geometry(name)
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        sim, _ = cases.bench_case(name)(mc, cl_devices=mc.cl.Context())
        n, t = cases.BENCH_RUN[name]
        rk = RefKernel(sim, geom, 'golden_bench_' + name)
        res = rk.run(n, t)
        out = {'packed_' + k: np.frombuffer(v, np.uint8) for k, v in rk.packed_bytes().items()}
        idx = np.flatnonzero(res['accu'])
        out.update(accu_idx=idx.astype(np.uint32), accu_val=res['accu'][idx],
                   accu_size=np.int64(res['accu'].size),
                   ints=res['ints'], floats=res['floats'],
                   rng_x_after=res['rng_x'][:t], lut=res['lut'],
                   num_kernels=res['num_kernels'], nphotons=n, nthreads=t,
                   rng_x0=sim.rng_seeds_x[:t], rng_a=sim.rng_seeds_a[:t])
        np.savez_compressed(os.path.join(HERE, 'bench_' + name + '.npz'), **out)
        print('bench', name, 'accu sum', int(res['accu'].sum()), 'nonzero bins', idx.size,
              'kernels', res['num_kernels'])


def make_traj(names):
    """tests/golden/traj_<case>.npz: per-packet trajectories of the reference
    kernel, ONE packet per work-item (n = t), for the north-star criterion
    "Trace trajectories within 1e-5 relative of the reference"."""
    ref_env.activate()
    import cases
    from refkernel import RefKernel
    import importlib
    for name in names or sorted(cases.TRAJ_RUN):
        if name in cases.BENCH_RUN:
            geom, make = cases.bench_geometry(name), cases.bench_case(name)
        else:
            geom, make = cases.GEOMETRY[name], cases.ALL_CASES[name]
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        sim, attrs = make(mc, cl_devices=mc.cl.Context())
        for k, v in attrs.items():
            setattr(sim, k, v)
        n = cases.TRAJ_RUN[name]
        rk = RefKernel(sim, geom, 'golden_' + name)
        res = rk.run(n, n)
        P = sim._packed['trace']
        ml = int(sim.trace.maxlen)
        _, do, co, _ = np.frombuffer(bytes(memoryview(P).cast('B')), np.uint32)[:4].tolist()
        rows = res['floats'][do:do + n*ml*8].reshape(n, ml, 8)
        cnt = res['ints'][co:co + n]
        idx = np.flatnonzero(res['accu'])
        np.savez_compressed(
            os.path.join(HERE, 'traj_' + name + '.npz'), rows=rows, counts=cnt,
            accu_idx=idx.astype(np.uint32), accu_val=res['accu'][idx],
            accu_size=np.int64(res['accu'].size), rng_x_after=res['rng_x'][:n],
            nphotons=n, maxlen=ml)
        print('traj', name, 'packets', n, 'mean events', float(cnt.mean()),
              'overflowed', int((cnt > ml).sum()))


def main(argv):
    if argv and argv[0] == 'bench':
        return make_bench(argv[1:])
    if argv and argv[0] == 'traj':
        return make_traj(argv[1:])
    ref_env.activate()
    if not os.path.exists(os.path.join(HERE, 'lut_hg08_2000.npz')):
        make_lut()
    import cases
    from refkernel import RefKernel
    import importlib
    names = argv or (list(cases.ALL_CASES) + list(cases.USER_CASES) + list(cases.DOUBLE_CASES))
    for name in names:
        geom = cases.GEOMETRY.get(name) or cases.USER_GEOMETRY.get(name) or \
            cases.DOUBLE_GEOMETRY[name]
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        make = cases.ALL_CASES.get(name) or cases.USER_CASES.get(name) or \
            cases.DOUBLE_CASES[name]
        sim, attrs = make(mc, cl_devices=mc.cl.Context())
        for k, v in attrs.items():
            setattr(sim, k, v)
        n, t = cases.GOLDEN_RUN.get(name) or cases.USER_RUN.get(name) or cases.DOUBLE_RUN[name]
        rk = RefKernel(sim, geom, 'golden_' + name)
        res = rk.run(n, t)
        out = {'packed_' + k: np.frombuffer(v, np.uint8) for k, v in rk.packed_bytes().items()}
        out.update(accu=res['accu'], ints=res['ints'], floats=res['floats'],
                   rng_x_after=res['rng_x'][:t], lut=res['lut'],
                   num_kernels=res['num_kernels'], nphotons=n, nthreads=t,
                   rng_x0=sim.rng_seeds_x[:t], rng_a=sim.rng_seeds_a[:t])
        if name in cases.SV_CASES or name in cases.SV_DOUBLE_CASES:
            # the reference's SamplingVolume kernel on the unfiltered trace rows
            tr = rk.packed_bytes()  # noqa: F841  (trace offsets read from the struct below)
            P = sim._packed['trace']
            ml = int(sim.trace.maxlen)
            _, do, co, _ = np.frombuffer(bytes(memoryview(P).cast('B')), np.uint32)[:4].tolist()
            rows = res["floats"][do:do + n*ml*8].copy()
            cnt = res["ints"][co:co + n].copy()
            svres = rk.sampling_volume(
                cases.make_sv(mc, cases.SV_DOUBLE_CASES.get(name, name)), cnt, rows)
            out.update(sv_accu=svres['accu'], sv_total_weight=svres['total_weight'],
                       sv_packed=np.frombuffer(svres['packed_sv'], np.uint8),
                       sv_packed_trace=np.frombuffer(svres['packed_sv_trace'], np.uint8))
        if geom == 'mcvox':
            out['voxels'] = np.ascontiguousarray(sim.voxels.data(sim)).view(np.int32)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'accu sum', int(res['accu'].sum()), 'kernels', res['num_kernels'],
              {k: v.size for k, v in out.items() if k.startswith('packed_')})


if __name__ == '__main__':
    main(sys.argv[1:])
