"""The reference's import statements and option objects work against this package
(host-side drop-in surface; xopto/mc{ml,vox,cyl}/__init__.py, mcbase/mcoptions.py)."""
import importlib

import pytest


@pytest.mark.parametrize('geom', ['mcml', 'mcvox', 'mccyl'])
def test_reference_style_imports(geom):
    pkg = importlib.import_module('pyxopto_b200.' + geom)
    for sub in ('mc', 'mcoptions', 'mcpf', 'mcfluence', 'mctrace', 'mctypes', 'mcsource',
                'mcdetector', 'mcsv', 'mcprogress', 'clinfo', 'mcutil'):
        assert importlib.import_module('pyxopto_b200.{}.{}'.format(geom, sub)) is \
            getattr(pkg, sub) or sub == 'mc'
    fib = importlib.import_module('pyxopto_b200.{}.mcutil.fiber'.format(geom))
    assert hasattr(fib, 'MultimodeFiber')
    mc = pkg.mc
    for name in ('mcoptions', 'mcpf', 'mcfluence', 'mctrace', 'mcsource', 'mcdetector',
                 'mctypes', 'mcsv', 'clinfo'):
        assert hasattr(mc, name), name


def test_user_dirs():
    import pyxopto_b200
    assert pyxopto_b200.USER_TMP_PATH and callable(pyxopto_b200.make_user_dirs)


def test_option_classes_of_the_reference_exist():
    """mcbase/mcoptions.py:119-732 + mcvox/mcoptions: same names, constructors and
    class-level instances; OpenCL-only options are accepted and ignored."""
    from pyxopto_b200.mcml import mcoptions as mo
    from pyxopto_b200.mcvox import mcoptions as vo
    for name in ('McOption', 'McBoolOption', 'McIntOption', 'McFloatOption', 'McTypeOption',
                 'McMethod', 'McUseFluenceCache', 'McUseHalfMath', 'McUseNativeMath',
                 'McIntLutMemory', 'McFloatLutMemory', 'McDebugMode', 'McUseEnhancedRng',
                 'McUseSoft64Atomics', 'McUseLottery', 'McMinimumPacketWeight',
                 'McPacketLotteryChance', 'McUsePackedStructures', 'McUseEvents'):
        assert hasattr(mo, name), name
        assert hasattr(vo, name), name
    assert vo.McMaterialMemory.constant_mem.cl_options == [('MC_MATERIAL_ARRAY_MEMORY', '__constant')]
    assert mo.McMethod.ar.cl_options == [('MC_METHOD', 1)]
    assert mo.McMethod(2).value == 2 and mo.McMethod('ar').value == 1
    with pytest.raises(ValueError):
        mo.McMethod(3)
    assert mo.McUseLottery.default is mo.McUseLottery.on
    assert mo.McUseNativeMath(True).cl_options == [('MC_USE_NATIVE_MATH', True)]
    assert mo.McBoolOption('MC_USE_LOTTERY', 0).value is False
    assert mo.McIntLutMemory('global').value == '__global'
    with pytest.raises(ValueError):
        mo.McFloatLutMemory('texture')
    assert mo.McOption.make_define('MC_METHOD', 1) == '#define MC_METHOD 1'
    assert mo.McOption.make_define('X', 0.5) == '#define X FP_LITERAL(0.5)'
    with pytest.raises(ValueError):
        mo.resolve_cl_options([mo.McMethod.ar], [mo.McMethod.aw])


def test_opencl_only_options_are_accepted_and_ignored():
    """A simulator built with every OpenCL-only option compiles to the same kernel
    source as one built without them."""
    import benchcfg
    from pyxopto_b200.mcvox import mc
    vo = mc.mcoptions
    plain = benchcfg.c3_vox(mc, n=21)
    opts = [vo.McMaterialMemory.constant_mem, vo.McFloatLutMemory.constant_mem,
            vo.McIntLutMemory('global'), vo.McUseNativeMath.on, vo.McUseHalfMath.off,
            vo.McUsePackedStructures.on, vo.McUseSoft64Atomics.on, vo.McUseFluenceCache.on,
            vo.McDebugMode.off]
    noisy = benchcfg.c3_vox(mc, n=21, options=opts)
    plain._pack(1000)
    noisy._pack(1000)
    assert plain.kernel_source() == noisy.kernel_source()
