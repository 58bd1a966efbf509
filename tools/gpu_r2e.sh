#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_full_trace_rows_are_consistent tests/test_gpu_validate.py::test_single_layer_uniform_fiber_trace tests/test_gpu_parity.py::test_throughput_mode_trace_statistics"
timeout 600 python -m pytest $T -m gpu -q -s 2>&1 | tail -12
echo "---- 128-bit stores"
XOPTO_TRACE_STORE=128 timeout 600 python -m pytest $T -m gpu -q -s 2>&1 | tail -12
timeout 300 python tools/probe_config.py c4_trace 1e6 2>&1 | sed -n 2,3p
timeout 300 python tools/probe_config.py c2_skin 1.25e8 2>&1 | sed -n 2,3p
timeout 900 python -m pytest tests -m gpu -q -k "aniso" 2>&1 | tail -5
