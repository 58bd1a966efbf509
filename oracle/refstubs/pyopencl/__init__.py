"""Import-time stand-in for ``pyopencl`` (TEST INFRASTRUCTURE, this container only).

The reference host layer (``/root/reference/xopto``) imports pyopencl at module
scope.  ``oracle/build_ref.py`` only needs the reference to *pack structs* and
*render kernel text*; no OpenCL call is ever made, so every name is inert.
"""


class _Bag:
    def __getattr__(self, name):
        return 0


class Device:
    pass


class Context:
    def __init__(self, devices=None, *a, **k):
        self.devices = list(devices or [])


class CommandQueue:
    def __init__(self, context=None, device=None, properties=None):
        self.context = context
        self.properties = 0


class Buffer:
    pass


class Program:
    pass


class Event:
    pass


class LocalMemory:
    pass


class RuntimeError(Exception):  # noqa: A001 - mirrors pyopencl.RuntimeError
    pass


mem_flags = _Bag()
command_queue_properties = _Bag()
device_type = _Bag()


def get_platforms():
    return []
