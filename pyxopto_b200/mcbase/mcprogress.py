"""Progress monitor (mirror of ``xopto/mcbase/mcprogress.py:32-270``).

A background thread reads the kernel's packet counter while ``Mc.run()`` is in
flight - on its **own stream** of the simulator's context, so the copy overtakes
the running kernel (the reference does the same with a second OpenCL queue,
mcprogress.py:72,226-246).  The counter the kernels advance is the reference's
``num_packets_done`` (mcml.template.c:460,790): in throughput mode warps claim it
32 packets at a time, in deterministic mode (static schedule) it is written once
at the end of the kernel.

    with ProgressMonitor(sim).start(nphotons) as monitor:
        sim.run(nphotons)

Subclass and override :py:meth:`ProgressMonitor.update`, or pass ``cb``, to
handle progress differently from the default one-line terminal bar.
"""
import shutil
import threading
import time

import numpy as np

from ..cu import abi


class ProgressMonitor:
    def __init__(self, mcsim, interval: float = 0.5, cb=None, cbargs=None, cbkwargs=None):
        self._mcsim = mcsim
        self._interval = float(interval)
        self._cb, self._cbargs, self._cbkwargs = cb, tuple(cbargs or ()), dict(cbkwargs or {})
        self._target = 0
        self._processed = 0
        self._threads = 0
        self._track = False
        self._stop = False
        self._terminate_on_stop = True
        self._stream = None
        self._wake = threading.Condition()
        self._thread = threading.Thread(target=self._proc, daemon=True)
        self._thread.start()

    # -- control (mcprogress.py:90-220) ------------------------------------------
    def start(self, target: int, terminate: bool = True) -> 'ProgressMonitor':
        if self._stop:
            raise RuntimeError('A terminated progress monitor can not be started!')
        with self._wake:
            self._target = int(target)
            self._processed = 0
            self._threads = 0
            self._terminate_on_stop = bool(terminate)
            self._track = True
            self._wake.notify_all()
        return self

    def resume(self, target: int = None):
        if self._stop:
            raise RuntimeError('A terminated progress monitor can not be resumed!')
        with self._wake:
            if target is not None:
                self._target = int(target)
            self._track = True
            self._wake.notify_all()

    def progress(self) -> float:
        return min(self._processed/max(self._target, 1), 1.0)

    def target(self) -> int:
        return self._target

    def processed(self) -> int:
        return self._processed

    def threads(self) -> int:
        return self._threads

    def stop(self):
        self._track = False
        self._processed = self._target
        if self._terminate_on_stop:
            self.terminate()

    def terminate(self):
        with self._wake:
            self._stop = True
            self._track = False
            self._wake.notify_all()
        if threading.current_thread() is not self._thread:
            self._thread.join(timeout=5.0)
        self._clear_line()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self._terminate_on_stop:
            self.terminate()
        else:
            self._track = False

    # -- polling thread ----------------------------------------------------------------
    def _poll(self) -> bool:
        """One read of (packets done, work-items finished); False if the simulator
        has no device buffers yet."""
        sim = self._mcsim
        ctx = getattr(sim, '_ctx', None)
        buf = getattr(sim, '_cl_buffers', {}).get(sim._counters_name()) if ctx is not None else None
        if buf is None:
            return False
        if self._stream is None or self._stream.ctx is not ctx:
            self._stream = abi.Stream(ctx)
        counters = np.zeros(2, dtype=np.uint32)
        buf.download(self._stream, counters)
        self._threads = int(counters[1])
        done = min(int(counters[0]), self._target)
        if done != self._processed:
            self._processed = done
            if self._cb is None:
                self.update()
            else:
                self._cb(self, *self._cbargs, **self._cbkwargs)
        return True

    def _proc(self):
        while True:
            with self._wake:
                while not self._stop and not (self._track and self._target > self._processed):
                    self._wake.wait(timeout=1.0)
                if self._stop:
                    return
            try:
                self._poll()
            except Exception:          # a monitor must never take the simulation down
                pass
            time.sleep(self._interval)

    # -- presentation --------------------------------------------------------------------
    def _clear_line(self):
        try:
            print(' '*shutil.get_terminal_size().columns, end='\r')
        except Exception:
            pass

    def update(self):
        """Called from the polling thread whenever the number of processed packets
        changed; override for custom handling (mcprogress.py:252-270)."""
        width = 40
        filled = int(round(self.progress()*width))
        print('|{}{}| {:5.1f}% {:,d}/{:,d}'.format(
            '#'*filled, '-'*(width - filled), 100.0*self.progress(),
            self._processed, self._target), end='\r')
