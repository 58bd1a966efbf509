"""Makes the read-only reference importable in THIS container (test infrastructure).

Only generator scripts (``oracle/build_ref.py``, ``tests/golden/make_golden.py``)
use this; nothing under ``pyxopto_b200`` and no ``-m gpu`` test imports it, because
``/root/reference`` does not exist on the GPU box.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get('XOPTO_REFERENCE', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'xopto'))


def activate():
    """Put the reference + import stubs on sys.path; returns the xopto package."""
    if not available():
        raise RuntimeError('reference checkout not found at ' + REFERENCE_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    stubs = os.path.join(here, 'refstubs')
    for p in (REFERENCE_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault(
        'PYXOPTO_USER_PATH', os.path.join(tempfile.gettempdir(), 'xopto_user'))
    os.makedirs(os.environ['PYXOPTO_USER_PATH'], exist_ok=True)
    import scipy.integrate as si
    if not hasattr(si, 'simps'):      # scipy >= 1.14 dropped the alias
        si.simps = si.simpson
    import xopto
    return xopto
