"""Flat read-write buffer allocators (mirror of ``xopto/mcbase/mcutil/buffer.py:121-287``).

Every plugin that needs device-side output asks the simulator for a slice of one
of three flat buffers (uint64 accumulators, floats, ints); offsets are assigned
in pack order and re-assigned on every run.
"""
import numpy as np


class BufferAllocation:
    def __init__(self, owner, offset: int, shape, dtype, download: bool = True):
        self.owner = owner
        self.offset = int(offset)
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.size = int(np.prod(self.shape))
        self.dtype = dtype
        self.download = download


class BufferAllocator:
    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)
        self.clear()

    def clear(self):
        self._allocations = []
        self._size = 0

    @property
    def size(self) -> int:
        return self._size

    def allocate(self, owner, shape, download: bool = True) -> BufferAllocation:
        alloc = BufferAllocation(owner, self._size, shape, self.dtype, download)
        self._allocations.append(alloc)
        self._size += alloc.size
        return alloc

    def allocations(self, owner=None):
        if owner is None:
            return list(self._allocations)
        return [a for a in self._allocations if a.owner is owner]

    def __len__(self):
        return len(self._allocations)
