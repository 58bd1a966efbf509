mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=240 -k "user" > gpurun_out/r05c_pytest_user.log 2>&1; tail -25 gpurun_out/r05c_pytest_user.log
