#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout=120 -k "user" 2>&1 | tail -6 | tee gpurun_out/r03o_pytest_user.log
