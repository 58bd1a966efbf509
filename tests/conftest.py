import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA GPU (run on the B200 box)')


def _has_gpu() -> bool:
    try:
        from pyxopto_b200.cu import abi
        return abi.device_count() > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


GPU_TEST_TIMEOUT_S = 300     # a kernel that never ends must not hold the GPU box


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        # (pytest-timeout, where installed: the watchdog thread ends the process, the
        # driver tears the context - and the runaway kernel - down)
        if config.pluginmanager.hasplugin('timeout'):
            for item in items:
                if 'gpu' in item.keywords and item.get_closest_marker('timeout') is None:
                    item.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method='thread'))
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
