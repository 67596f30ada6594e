timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2w_pytest.log 2>&1; tail -4 gpurun_out/r2w_pytest.log
