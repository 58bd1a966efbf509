#!/bin/bash
mkdir -p gpurun_out
T=r02z
timeout 420 python -m pytest tests -m gpu -q -x --timeout=90 -k "double or both_mcvox or user or rayleigh" > gpurun_out/${T}_pytest_sel.log 2>&1; tail -15 gpurun_out/${T}_pytest_sel.log
