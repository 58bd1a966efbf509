"""Generates tests/golden/*.npz with the REFERENCE itself (this container only).

For every case in tests/cases.py the reference package packs its structs and
renders its kernel; the kernel text is compiled unchanged for the CPU
(oracle/refkernel.py) and run under the static block schedule.  Stored per case:
the raw packed struct bytes, the float LUT pool, and the resulting accumulator /
int / float buffers + advanced RNG states.  These pin (a) the host mirror's
packing and (b) the oracle restatement on machines without /root/reference.

    python tests/golden/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import ref_env  # noqa: E402


def make_lut():
    ref_env.activate()
    from xopto import pf
    params, lut = pf.Hg(0.8).mclut(2000)
    np.savez_compressed(os.path.join(HERE, 'lut_hg08_2000.npz'),
                        params=np.asarray(params, np.float64),
                        lut=np.asarray(lut, np.float64))


def main(argv):
    ref_env.activate()
    if not os.path.exists(os.path.join(HERE, 'lut_hg08_2000.npz')):
        make_lut()
    import cases
    from refkernel import RefKernel
    import importlib
    names = argv or (list(cases.ALL_CASES) + list(cases.USER_CASES))
    for name in names:
        geom = cases.GEOMETRY.get(name) or cases.USER_GEOMETRY[name]
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        make = cases.ALL_CASES.get(name) or cases.USER_CASES[name]
        sim, attrs = make(mc, cl_devices=mc.cl.Context())
        for k, v in attrs.items():
            setattr(sim, k, v)
        n, t = cases.GOLDEN_RUN.get(name) or cases.USER_RUN[name]
        rk = RefKernel(sim, geom, 'golden_' + name)
        res = rk.run(n, t)
        out = {'packed_' + k: np.frombuffer(v, np.uint8) for k, v in rk.packed_bytes().items()}
        out.update(accu=res['accu'], ints=res['ints'], floats=res['floats'],
                   rng_x_after=res['rng_x'][:t], lut=res['lut'],
                   num_kernels=res['num_kernels'], nphotons=n, nthreads=t,
                   rng_x0=sim.rng_seeds_x[:t], rng_a=sim.rng_seeds_a[:t])
        if name in cases.SV_CASES:
            # the reference's SamplingVolume kernel on the unfiltered trace rows
            tr = rk.packed_bytes()  # noqa: F841  (trace offsets read from the struct below)
            P = sim._packed['trace']
            ml = int(sim.trace.maxlen)
            _, do, co, _ = np.frombuffer(bytes(memoryview(P).cast('B')), np.uint32)[:4].tolist()
            rows = res["floats"][do:do + n*ml*8].copy()
            cnt = res["ints"][co:co + n].copy()
            svres = rk.sampling_volume(cases.make_sv(mc, name), cnt, rows)
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
