"""Batch runners - callers of ``Mc.run`` that repeat a simulation until a detector
has collected enough signal.  Mirror of ``xopto/mcbase/mcrun.py:28-263`` (bases) and
of the per-geometry modules ``xopto/mcml/mcrun/mcrun.py:32-257``,
``xopto/mcvox/mcrun/mcrun.py:32-257``, ``xopto/mccyl/mcrun/mcrun.py:32-237``: same
class names, constructor arguments, ``run(mc_obj, min_packets, out, *args, **kwargs)``
call, ``process_batch`` hook and properties.

One loop serves both criteria.  Each batch is ``mc_obj.run(batch_size, out=out, ...)``:
the accumulators of all batches are summed by the engine (on the device when
``Mc.lazy_fluence`` is set - only the small detector arrays the stop test reads come
to the host between two batches; the fluence grid is downloaded once, when the
caller reads it).

Reference quirk, not reproduced: ``RunMinPacketsBase.run`` of the reference tests
``self.min_weight``, an attribute that class never sets (``mcrun.py:196``), so
``RunMinPacketsTrace`` raises ``AttributeError`` there on its first batch.  Here it
does what its documentation says: it stops once ``process_batch`` reports at least
``min_packets`` collected packets.
"""
import sys
import types
from typing import Tuple

import numpy as np


class _BatchRunner:
    """The loop shared by the weight and the packet-count criterion."""

    def __init__(self, batch_size: int, selection=None):
        self._selection = slice(None) if selection is None else selection
        self._batch_size = int(batch_size)
        if self._batch_size <= 0:
            raise ValueError('The batch size must be a positive integer!')
        self._num_packets = 0
        self._total_weight = 0.0
        self._collected = 0

    def _satisfied(self, total_weight: float, num_collected) -> bool:
        raise NotImplementedError

    def run(self, mc_obj, min_packets: int = None, out: tuple = None,
            *args, **kwargs) -> tuple:
        """Simulate batches of ``batch_size`` packets until the criterion of the
        class holds and at least ``min_packets`` packets were launched; at least
        one batch runs when ``out`` is None (``mcrun.py:49-92``)."""
        min_launched = 0 if min_packets is None else int(min_packets)
        launched, weight, collected = 0, 0.0, 0
        while True:
            if out is not None and launched >= min_launched and \
                    self._satisfied(weight, collected):
                break
            out = mc_obj.run(self._batch_size, *args, out=out, **kwargs)
            weight, n, out = self.process_batch(out)
            weight = float(weight)
            collected = collected if n is None else int(n)
            launched += self._batch_size
        self._num_packets, self._total_weight, self._collected = launched, weight, collected
        return out

    def process_batch(self, result: tuple) -> Tuple[float, int, tuple]:
        """Hook of the subclasses: (total weight, collected packets or None, result
        to hand to the next batch)."""
        return 0.0, 0, result

    selection = property(lambda self: self._selection, None, None,
                         'Slice / index of the detector data that counts towards the weight.')
    batch_size = property(lambda self: self._batch_size, None, None,
                          'Packets launched per batch.')
    n = property(lambda self: self._num_packets, None, None,
                 'Packets launched by the last call of run().')
    weight = property(lambda self: self._total_weight, None, None,
                      'Weight collected at the end of the last call of run().')


class RunMinWeightBase(_BatchRunner):
    """Repeat until the selected detector holds at least ``min_weight``
    (``mcrun.py:28-143``)."""

    def __init__(self, min_weight: float, batch_size: int, selection=None):
        super().__init__(batch_size, selection)
        self._min_weight = float(min_weight)

    def _satisfied(self, total_weight, num_collected):
        return total_weight >= self._min_weight

    min_weight = property(lambda self: self._min_weight, None, None,
                          'Weight the detector has to collect.')


class RunMinPacketsBase(_BatchRunner):
    """Repeat until at least ``min_packets`` packets were collected
    (``mcrun.py:146-263``; see the module header for the reference's slip)."""

    def __init__(self, min_packets: int, batch_size: int, selection=None):
        super().__init__(batch_size, selection)
        self._min_packets = int(min_packets)

    def _satisfied(self, total_weight, num_collected):
        return num_collected >= self._min_packets

    min_packets = property(lambda self: self._min_packets, None, None,
                           'Packets that have to be collected.')


def _detector_weight_runner(mc_module, location: str, name: str):
    """Class ``RunMinWeight<Location>`` of one geometry: the weight is the sum of
    the selected raw accumulator entries of the detector at ``location``
    (``mcml/mcrun/mcrun.py:32-160``)."""
    default_type = mc_module.mcdetector.DetectorDefault

    class _Runner(RunMinWeightBase):
        def __init__(self, min_weight: float, batch_size: int, selection=None):
            super().__init__(min_weight, batch_size, selection)

        def process_batch(self, result):
            detectors = result[2]
            detector = None if detectors is None else getattr(detectors, location)
            if detector is None or type(detector) is default_type:
                raise RuntimeError(
                    'This MC simulator instance does not use the configured'
                    ' ("{}") detector!'.format(location))
            raw = detector.raw
            if self._selection is not None:
                raw = raw[self._selection]
            return np.sum(raw), None, result

        location = property(lambda self: location, None, None, 'Detector location.')

    _Runner.__name__ = _Runner.__qualname__ = name
    _Runner.__doc__ = 'Batches until the {} detector has collected min_weight.'.format(location)
    return _Runner


def _trace_process_batch(self, result):
    trace = result[0]
    if trace is None:
        raise RuntimeError('This MC simulator instance does not use trace!')
    return np.sum(trace.terminal['w']), len(trace), result


def geometry_module(package: str, mc_module, locations) -> types.ModuleType:
    """Build ``<package>.mcrun`` for one geometry (top / bottom / specular for the
    layered and the voxel simulator, outer / specular for the cylindrical one) and
    register it under the reference's import path (``from xopto.mcml.mcrun import
    RunMinWeightTop`` style)."""
    mod = types.ModuleType(package + '.mcrun', __doc__)
    for location in locations:
        name = 'RunMinWeight' + location.capitalize()
        cls = _detector_weight_runner(mc_module, location, name)
        cls.__module__ = mod.__name__
        setattr(mod, name, cls)

    class RunMinWeightTrace(RunMinWeightBase):
        """Batches until the traced packets carry min_weight at their last event
        (``mcml/mcrun/mcrun.py:163-208``)."""
        def __init__(self, min_weight: float, batch_size: int):
            super().__init__(min_weight, batch_size, None)
        process_batch = _trace_process_batch

    class RunMinPacketsTrace(RunMinPacketsBase):
        """Batches until the trace holds min_packets packets
        (``mcml/mcrun/mcrun.py:211-257``)."""
        def __init__(self, min_packets: int, batch_size: int):
            super().__init__(min_packets, batch_size, None)
        process_batch = _trace_process_batch

    mod.RunMinWeightTrace, mod.RunMinPacketsTrace = RunMinWeightTrace, RunMinPacketsTrace
    mod.RunMinWeightBase, mod.RunMinPacketsBase = RunMinWeightBase, RunMinPacketsBase
    for cls in (RunMinWeightTrace, RunMinPacketsTrace):
        cls.__module__ = mod.__name__
    mod.mcrun = mod                      # xopto.<geometry>.mcrun.mcrun is importable as well
    sys.modules[mod.__name__] = mod
    sys.modules.setdefault(mod.__name__ + '.mcrun', mod)
    return mod
