mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_mcrun.py -m gpu -q --timeout=240 > gpurun_out/r05a_pytest_mcrun.log 2>&1; tail -5 gpurun_out/r05a_pytest_mcrun.log
