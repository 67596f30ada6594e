timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; tail -6 gpurun_out/r2s_pytest.log
python tools/quick_rate.py 65536 40 60 2>&1 | tail -1
timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1
