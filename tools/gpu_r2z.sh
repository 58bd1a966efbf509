#!/bin/bash
mkdir -p gpurun_out
T=r02z
timeout 900 python -m pytest tests -m gpu -q -x -k "double or both_mcvox or user" > gpurun_out/${T}_pytest_sel.log 2>&1; tail -15 gpurun_out/${T}_pytest_sel.log
