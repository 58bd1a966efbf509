"""CUDA device discovery (mirror of the keyword lookup in ``xopto/cl/clinfo.py``).

A device is identified by its CUDA ordinal; ``Device`` objects are what the
``cl_devices`` argument of ``Mc`` accepts.
"""
from typing import List

from ..cu import abi


class Device:
    def __init__(self, ordinal: int):
        self.ordinal = int(ordinal)
        self.info = abi.device_info(self.ordinal)

    @property
    def name(self) -> str:
        return self.info['name']

    def __repr__(self):
        return 'Device({}: {})'.format(self.ordinal, self.name)


def gpus() -> List[Device]:
    """All CUDA devices (raises RuntimeError when no driver/GPU is present)."""
    return [Device(i) for i in range(abi.device_count())]


def gpu(index: int = 0) -> Device:
    devs = gpus()
    if index >= len(devs):
        raise RuntimeError('CUDA device {} not available!'.format(index))
    return devs[index]


def device(what=None, index: int = 0) -> Device:
    """First (index-th) device whose name contains any of the keywords."""
    if what is None:
        return gpu(index)
    if isinstance(what, str):
        what = [what]
    found = [d for d in gpus()
             if any(k.lower() in d.name.lower() for k in what)]
    if index >= len(found):
        raise RuntimeError('No CUDA device matches {}!'.format(what))
    return found[index]


def info(dev: Device = None) -> str:
    dev = dev or gpu()
    return '\n'.join('{}: {}'.format(k, v) for k, v in dev.info.items())
