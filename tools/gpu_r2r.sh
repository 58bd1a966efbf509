#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_config.py c5_cyl 1e7 2>&1 | sed -n 2,3p
timeout 900 python -m pytest tests -m gpu -q -k "mccyl or c5_cyl or cyl" 2>&1 | tail -4
