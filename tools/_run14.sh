export COSIM_PRINT_OCC=1
python tools/pool_check.py 8192 40 2>&1 | tail -4
python tools/pool_check.py 8192 12 w4_p_v2 stairs_up_hard 2>&1 | tail -2
python tools/pool_check.py 8192 20 humanoid_p_v0 slope_hard 2>&1 | tail -2
for r in 0 2 3 4 6; do echo "R=$r"; COSIM_POOL_R=$r python tools/quick_rate.py 65536 20 5 2>&1 | tail -1; done
for b in 28 20 22 62; do echo "bounds=$b"; COSIM_POOL_BOUNDS=$b python tools/quick_rate.py 65536 20 5 2>&1 | tail -1; done
timeout 300 python tools/soak.py 150 65536 2>&1 | tail -2
timeout 300 python tools/soak.py 60 16384 w4_p_v2 stairs_up_hard 2>&1 | tail -2
timeout 300 python tools/soak.py 100 16384 humanoid_p_v0 slope_hard 2>&1 | tail -2
