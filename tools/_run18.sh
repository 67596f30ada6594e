timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; tail -12 gpurun_out/r2o_pytest.log
python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
timeout 300 python tools/soak.py 100 65536 2>&1 | tail -1
