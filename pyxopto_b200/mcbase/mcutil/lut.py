"""Read-only lookup-table pool (mirror of ``xopto/mcbase/mcutil/lut.py:408-578``):
tables are de-duplicated by value and concatenated into one float array."""
import numpy as np


class LutEntry:
    def __init__(self, manager, data: np.ndarray, offset: int):
        self.manager = manager
        self.data = data
        self.offset = int(offset)
        self.size = int(data.size)


class LutManager:
    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.clear()

    def clear(self):
        self._entries = []
        self._size = 0

    def append(self, data: np.ndarray, force: bool = False) -> LutEntry:
        flat = np.asarray(data).ravel()
        if not force:
            for e in self._entries:
                if e.data.shape == flat.shape and np.array_equal(e.data, flat):
                    return e
        entry = LutEntry(self, flat, self._size)
        self._entries.append(entry)
        self._size += flat.size
        return entry

    def pack_into(self, target: np.ndarray = None) -> np.ndarray:
        if target is None or target.size < self._size:
            target = np.zeros(max(self._size, 1), dtype=self.dtype)
        for e in self._entries:
            target[e.offset:e.offset + e.size] = e.data
        return target

    @property
    def size(self) -> int:
        return self._size

    def __len__(self):
        return len(self._entries)
