export COSIM_PRINT_OCC=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; tail -15 gpurun_out/r2j_pytest.log
python tools/quick_rate.py 65536 20 5 > gpurun_out/r2j_quick.log 2>&1; tail -1 gpurun_out/r2j_quick.log
timeout 300 python tools/soak.py 200 65536 > gpurun_out/r2j_soak_bench.log 2>&1; tail -4 gpurun_out/r2j_soak_bench.log
timeout 300 python tools/soak.py 60 16384 w4_p_v2 stairs_up_hard > gpurun_out/r2j_soak_w4.log 2>&1; tail -2 gpurun_out/r2j_soak_w4.log
timeout 300 python tools/soak.py 100 16384 humanoid_p_v0 slope_hard > gpurun_out/r2j_soak_hum.log 2>&1; tail -2 gpurun_out/r2j_soak_hum.log
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -3 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
