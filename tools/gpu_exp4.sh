#!/bin/bash
timeout 300 python tools/exp_knobs.py c4_trace 1e6 "refill_lanes=3;refill_lanes=3,min_blocks=3;refill_lanes=2,min_blocks=3;refill_lanes=3,min_blocks=2;refill_lanes=1,min_blocks=3" 2>&1 | tail -6
