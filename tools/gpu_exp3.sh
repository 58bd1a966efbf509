#!/bin/bash
timeout 300 python tools/exp_knobs.py c2_skin 4e7 "refill_lanes=8,fluence_window_aspect=0.5;refill_lanes=8,fluence_window_aspect=0.7;refill_lanes=8,fluence_window_aspect=0.85;refill_lanes=8,fluence_window_aspect=1.0;refill_lanes=8,fluence_window_aspect=1.2;refill_lanes=10,fluence_window_aspect=0.85;refill_lanes=6,fluence_window_aspect=0.85" 2>&1 | tail -9
