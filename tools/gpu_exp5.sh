#!/bin/bash
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/exp_knobs.py c2_skin 1.25e8 "refill_lanes=8" 2>&1 | tail -1
timeout 300 python tools/exp_knobs.py c1_slab 1e7 "refill_lanes=3" 2>&1 | tail -1
timeout 300 python tools/exp_knobs.py c3_vox 1e8 "wait_lanes=16" 2>&1 | tail -1
timeout 300 python tools/exp_knobs.py c5_cyl 1e7 "refill_lanes=1" 2>&1 | tail -1
