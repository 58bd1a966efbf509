"""pyxopto_b200 - B200-native photon-packet Monte Carlo engine behind the
PyXOpto ``Mc`` API (layered ``mcml``, voxelised ``mcvox``, cylindrical ``mccyl``).

The package mirrors the reference's user-facing interface for the hot path
(``Mc(...)``, ``run()``, plugin objects and their packed structs; see
``xopto/mc{ml,vox,cyl}/mc.py``) on top of ``libxopto_b200.so`` (C ABI over the
CUDA driver API + NVRTC, ``include/xopto_b200.h``) and hand-written CUDA kernels
for sm_100a (``csrc/kernels``).  There is no CPU fallback.
"""
import os

__version__ = '0.1.0'

ROOT_PATH = os.path.dirname(os.path.abspath(__file__))
KERNEL_PATH = os.path.join(ROOT_PATH, 'csrc', 'kernels')
DATA_PATH = os.path.join(ROOT_PATH, 'data')
KERNEL_CACHE_PATH = os.environ.get(
    'PYXOPTO_B200_KCACHE', os.path.join(ROOT_PATH, '_kcache'))
VERBOSE = bool(int(os.environ.get('PYXOPTO_VERBOSE', '0')))

# user directories of the reference (xopto/__init__.py:58-140): scripts write
# exported kernel sources / temporary files there
USER_PATH = os.environ.get('PYXOPTO_USER_PATH', os.path.join(os.path.expanduser('~'), '.xopto'))
USER_DATA_PATH = os.path.join(USER_PATH, 'data')
USER_BIN_PATH = os.path.join(USER_PATH, 'bin')
USER_TMP_PATH = os.path.join(USER_PATH, 'tmp')


def make_user_dirs():
    """Creates the user directories (xopto/__init__.py:128)."""
    for path in (USER_DATA_PATH, USER_BIN_PATH, USER_TMP_PATH):
        os.makedirs(path, exist_ok=True)
