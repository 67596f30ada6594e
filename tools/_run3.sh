timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
for k in 5 16 32 48; do echo "K=$k"; COSIM_CN_K=$k COSIM_PRINT_OCC=1 timeout 300 python tools/soak.py 60 16384 w4_p_v2 stairs_up_hard 2>&1 | tail -3; done
timeout 300 python tools/soak.py 100 16384 humanoid_p_v0 slope_hard 2>&1 | tail -2
python tools/phase_profile.py w4_p_v2 stairs_up_hard 4096 3 2>&1 | grep -v histogram
python tools/phase_profile.py flamingo_p_v3 rocky_hard 16384 10 2>&1 | grep -v histogram
