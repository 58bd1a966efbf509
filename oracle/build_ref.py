"""Builds oracle/_ref/libref_<config>.so: the REFERENCE's own kernel for the
bench configurations, rendered by the reference's Python and compiled unchanged
with gcc (TEST / BASELINE INFRASTRUCTURE; only runs where /root/reference exists).

The .so files are git-ignored but travel to the GPU box, where ``bench.py`` uses
them as the CPU baseline of kind "reference".
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_env  # noqa: E402


def main():
    if not ref_env.available():
        print('reference not present; nothing to build')
        return 0
    ref_env.activate()
    import benchcfg
    import refkernel
    for name, make in benchcfg.CONFIGS.items():
        geom = benchcfg.GEOMETRY[name]
        mc = importlib.import_module('xopto.{}.mc'.format(geom))
        sim = make(mc, cl_devices=mc.cl.Context())
        sim._pack(1000)
        sim._build_src()
        so = refkernel.compile_rendered(
            sim._cl_src, geom, name, cflags=['-O3', '-ffp-contract=off'], stable=True)
        print('built', so)
    return 0


if __name__ == '__main__':
    sys.exit(main())
