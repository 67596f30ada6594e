set -x
export COSIM_PRINT_OCC=1
python tools/quick_rate.py 65536 20 5 > gpurun_out/r2a_quick.log 2>&1
tail -3 gpurun_out/r2a_quick.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python tools/soak.py 100 16384 w4_p_v2 stairs_up_hard > gpurun_out/r2a_soak_w4.log 2>&1
tail -4 gpurun_out/r2a_soak_w4.log
timeout 300 python tools/soak.py 100 16384 humanoid_p_v0 slope_hard > gpurun_out/r2a_soak_hum.log 2>&1
tail -4 gpurun_out/r2a_soak_hum.log
timeout 300 python tools/soak.py 150 65536 > gpurun_out/r2a_soak_bench.log 2>&1
tail -4 gpurun_out/r2a_soak_bench.log
