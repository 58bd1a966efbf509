#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=120 -k "trace_statistics" 2>&1 | grep -v "^  \|Warning" | tail -40 | cut -c1-250 | tee gpurun_out/r03k_pytest_sel.log
