#!/bin/bash
timeout 1500 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "mcml or c1 or c2 or staged or trace or user or fast_mode" 2>&1 | tail -4
timeout 300 python tools/exp_knobs.py c2_skin 1.25e8 "refill_lanes=8;refill_lanes=7;refill_lanes=9" 2>&1 | tail -3
timeout 300 python tools/exp_knobs.py c1_slab 1e7 "refill_lanes=3;refill_lanes=2" 2>&1 | tail -2
timeout 300 python tools/exp_knobs.py c4_trace 1e6 "refill_lanes=3" 2>&1 | tail -1
