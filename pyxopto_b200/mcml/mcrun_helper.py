"""``McRunHelper`` - scaffold of the reference's dataset run scripts
(``xopto/mcml/mcrun/helper.py:32-686``): a subclass overrides the ``create_*`` hooks
(one simulator per script) and the ``update_*`` hooks (called before every sample),
``run_one`` / ``run_batch`` return dictionaries of configuration + results ready for
``np.savez_compressed``.

Same class name, hooks, call signatures and result keys.  What differs is below the
API: ``run_batch`` streams its samples through the sweep driver (``mcsweep.Sweep``:
one compiled kernel, packed structs rewritten per sample, accumulators double-buffered
on the device, results crossing PCIe while the next sample runs) whenever the
simulator records no trace and every sample launches the same number of packets;
otherwise it is the reference's loop of ``run_one`` calls.
"""
import argparse
import os
import time
from typing import Callable, Dict, List

import numpy as np

from . import mc
from .. import mcsweep


class McRunHelper:
    @staticmethod
    def cli_input(first: int = 0, n: int = 1000, batch: int = 1000, packets: int = 100e6,
                  device: str = None, device_index: int = 0, root_dir: str = None,
                  mc_dir: str = None, processed_dir: str = None, verbose: bool = False,
                  label: str = None, argv=None) -> dict:
        """Command-line options of a run script (helper.py:34-136); ``device`` is the
        unique part of the GPU's name, ``device_index`` the CUDA device ordinal.
        ``argv``: parse this list instead of ``sys.argv[1:]``."""
        parser = argparse.ArgumentParser(
            description=label or 'Command line arguments for MC simulations')
        parser.add_argument('-d', '--device', dest='device', type=str, default=device or '',
                            help='Unique part of the device name.')
        parser.add_argument('-i', '--index', dest='device_index', type=int,
                            default=int(device_index), help='Device index.')
        parser.add_argument('-f', '--first', dest='first', type=int, default=int(first),
                            help='Zero-based index of the first simulated sample.')
        parser.add_argument('-n', '--number', dest='number', type=int, default=int(n),
                            help='Number of samples that will be simulated.')
        parser.add_argument('-b', '--batch', dest='batch', type=int, default=int(batch),
                            help='Number of samples saved into a single file.')
        parser.add_argument('-p', '--packets', dest='packets', type=float,
                            default=int(packets), help='Packets launched per simulation.')
        parser.add_argument('--root-dir', dest='root_dir', type=str,
                            default=str(root_dir or os.getcwd()),
                            help='Root directory of the dataset.')
        parser.add_argument('--output-dir', dest='mc_dir', type=str,
                            default=str(mc_dir or 'mc'),
                            help='Output subdirectory for the MC simulations.')
        parser.add_argument('--processed-dir', dest='processed_dir', type=str,
                            default=str(processed_dir or 'processed'),
                            help='Output subdirectory for the postprocessed data.')
        parser.add_argument('-v', '--verbose', dest='verbose', action='store_true',
                            default=bool(verbose), help='Turn on verbose mode.')
        a = parser.parse_args(argv)
        return {'first': a.first, 'n': a.number, 'batch': a.batch, 'packets': int(a.packets),
                'root_dir': a.root_dir, 'mc_dir': a.mc_dir, 'processed_dir': a.processed_dir,
                'device': a.device if a.device else None, 'device_index': a.device_index,
                'verbose': a.verbose}

    def __init__(self, *args, **kwargs):
        """Arguments go to the simulator's constructor after layers, source and
        detectors (helper.py:138-156)."""
        self._mc_obj = self.create_mc(*args, **kwargs)
        self._sweep = None

    # -- construction hooks (helper.py:158-318) ---------------------------------------
    def create_layers(self):
        # (the reference's default reads ``mc.mclyer`` and raises AttributeError:
        #  helper.py:172; here the three void layers it meant to build)
        void = dict(d=float('inf'), mua=0.0, mus=0.0, n=1.0)
        return mc.mclayer.Layers([mc.mclayer.Layer(pf=mc.mcpf.Hg(0.8), **void)
                                  for _ in range(3)])

    def create_source(self):
        return mc.mcsource.Line()

    def create_surface(self):
        return None

    def create_detectors(self):
        return mc.mcdetector.Detectors(
            top=mc.mcdetector.Radial(mc.mcdetector.Axis(0.0, 5.0, 500)),
            bottom=mc.mcdetector.Radial(mc.mcdetector.Axis(0.0, 5.0, 500)),
            specular=mc.mcdetector.Total())

    def create_fluence(self):
        return None

    def create_trace(self):
        return None

    def create_mc(self, *args, **kwargs):
        source, surface = self.create_source(), self.create_surface()
        layers, detectors = self.create_layers(), self.create_detectors()
        fluence, trace = self.create_fluence(), self.create_trace()
        return mc.Mc(layers, source, detectors, *args,
                     trace=trace, fluence=fluence, surface=surface, **kwargs)

    # -- per-sample hooks (helper.py:320-428) -----------------------------------------
    def update_layers(self):
        pass

    def update_source(self):
        pass

    def update_surface(self):
        pass

    def update_detectors(self):
        pass

    def update_fluence(self):
        pass

    def update_trace(self):
        pass

    def update_rmax(self):
        pass

    def verbose_print(self):
        pass

    def update(self):
        for hook in (self.update_layers, self.update_source, self.update_surface,
                     self.update_detectors, self.update_fluence, self.update_trace,
                     self.update_rmax):
            hook()

    mc_obj = property(lambda self: self._mc_obj, None, None, 'Monte Carlo simulator instance')

    # -- results (helper.py:434-543) ----------------------------------------------------
    def collect_mc_config(self) -> dict:
        sim = self._mc_obj

        def d(obj):
            return None if obj is None else obj.todict()
        return {'source': d(sim.source), 'surface': d(sim.surface), 'layers': d(sim.layers),
                'detectors': d(sim.detectors), 'fluence': d(sim.fluence),
                'trace': d(sim.trace), 'rmax': sim.rmax, 'run_report': dict(sim.run_report)}

    def collect_detectors(self, result) -> dict:
        data = {'reflectance': None, 'transmittance': None, 'specular': None}
        if result is not None:
            for key, det in (('reflectance', result.top), ('transmittance', result.bottom),
                             ('specular', result.specular)):
                if type(det) is not mc.mcdetector.DetectorDefault:
                    data[key] = det.reflectance
        return data

    def collect_fluence(self, result) -> dict:
        return {'data': None if result is None else result.data}

    def collect_trace(self, result):
        return None if result is None else {'data': result.data, 'n': result.n}

    def collect_one(self, nphotons: int, trace_res=None, fluence_res=None,
                    detectors_res=None) -> dict:
        return {'epoch_timestamp': time.time(), 'mc': self.collect_mc_config(),
                'num_packets': nphotons,
                'detectors': self.collect_detectors(detectors_res),
                'fluence': self.collect_fluence(fluence_res),
                'trace': self.collect_trace(trace_res)}

    # -- runs (helper.py:545-686) -------------------------------------------------------
    def run_one(self, nphotons, verbose: bool = False, *args, **kwargs) -> dict:
        if not isinstance(nphotons, (int, float, np.integer)):
            nphotons = nphotons(self)
        self.update()
        if verbose:
            self.verbose_print()
        return self.collect_one(nphotons, *self._mc_obj.run(nphotons, *args, **kwargs))

    streamed = True      # run_batch may go through the sweep driver

    def run_batch(self, size: int, nphotons, first: int = 0, verbose: bool = False,
                  *args, **kwargs) -> List[Dict]:
        sim = self._mc_obj
        fixed = isinstance(nphotons, (int, float, np.integer))
        if not (self.streamed and fixed and sim.trace is None and not args
                and set(kwargs) <= {'wgsize', 'maxthreads'}):
            return self._run_batch_sequential(size, nphotons, first, verbose, *args, **kwargs)
        # One pass of the sweep driver: `apply` is this helper's update() - it runs
        # while the previous sample is still on the GPU - and records the sample's
        # configuration at that moment, as collect_one would.
        nphotons = int(nphotons)
        configs, stamps = {}, {}

        def apply(_sim, i):
            self.update()
            if verbose:
                self.verbose_print()
            configs[i], stamps[i] = self.collect_mc_config(), time.time()
        if self._sweep is None:
            self._sweep = mcsweep.Sweep(sim)
        t0 = time.perf_counter()
        indices, rows = self._sweep.run(list(range(size)), nphotons, apply=apply, **kwargs)
        batch = []
        for i, row in zip(indices, rows):
            _, flu, det = mcsweep.results_from_row(sim, row, nphotons)
            data = {'epoch_timestamp': stamps[int(i)], 'mc': configs[int(i)],
                    'num_packets': nphotons, 'detectors': self.collect_detectors(det),
                    'fluence': self.collect_fluence(flu), 'trace': None,
                    'index': first + int(i)}
            batch.append(data)
        if verbose:
            dt = time.perf_counter() - t0
            print('{}-{}: {} samples - {:.1f} samples/h'.format(
                first, first + size, size, size/dt*3600 if dt > 0 else float('nan')))
        return batch

    def _run_batch_sequential(self, size, nphotons, first, verbose, *args, **kwargs):
        batch, t0, rate = [], time.perf_counter(), float('nan')
        for i in range(size):
            if verbose:
                print('{}-{}: {}/{} - {:.1f} samples/h'.format(
                    first, first + size, i + 1, size, rate))
            data = self.run_one(nphotons, verbose, *args, **kwargs)
            data['index'] = first + i
            batch.append(data)
            rate = (i + 1)/(time.perf_counter() - t0)*3600
        return batch
