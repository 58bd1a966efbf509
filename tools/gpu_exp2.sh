#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "throughput or scale or deterministic_mode" 2>&1 | tail -3
timeout 300 python tools/exp_knobs.py c2_skin 4e7 "refill_lanes=8;refill_lanes=7;refill_lanes=6;refill_lanes=8,fluence_window_aspect=1.0;refill_lanes=8,fluence_window_aspect=1.5;refill_lanes=8,fluence_window_aspect=2.5;refill_lanes=8,fluence_window_aspect=3.0;refill_lanes=8,fluence_window_aspect=4.0" 2>&1 | tail -9
