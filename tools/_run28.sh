timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1; tail -4 gpurun_out/r2v_pytest.log
python tools/quick_rate.py 65536 40 60 2>&1 | tail -1
python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
COSIM_SOAK_BLOCK=30 timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1 | cut -c1-60
COSIM_SOAK_BLOCK=15 timeout 300 python tools/soak.py 30 8192 w4_p_v2 stairs_up_hard 2>&1 | tail -1 | cut -c1-60
COSIM_SOAK_BLOCK=30 timeout 300 python tools/soak.py 60 65536 flamingo_light_v1 flat 2>&1 | tail -1 | cut -c1-60
python tools/pool_check.py 2>&1 | tail -2
