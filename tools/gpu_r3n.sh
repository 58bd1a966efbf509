#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout=120 -k "both_mcvox or vox or host_logic" 2>&1 | tail -6 | tee gpurun_out/r03n_pytest_sel.log
