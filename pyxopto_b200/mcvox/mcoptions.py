"""Kernel options of the voxelised simulator (mirror of
``xopto/mcvox/mcoptions/__init__.py``): the shared options plus
``McMaterialMemory``."""
from ..mcbase.mcoptions import *                    # noqa: F401,F403
from ..mcbase.mcoptions import _memory_option, _named

McMaterialMemory = _named(_memory_option(
    'MC_MATERIAL_ARRAY_MEMORY', 'Material data', 'constant',
    'OpenCL memory space of the material array (mcvox/mcoptions/__init__.py:25); '
    'ignored: the material table is staged in shared memory.'), 'McMaterialMemory')
