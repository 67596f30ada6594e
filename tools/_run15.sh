python tools/pool_check.py 8192 20 2>&1 | tail -1
COSIM_POOL_MODE=1 python tools/pool_check.py 8192 20 2>&1 | tail -1
for r in 2 3 5; do echo "flamingo mode1 R=$r"; COSIM_POOL_MODE=1 COSIM_POOL_R=$r python tools/soak.py 100 65536 2>&1 | tail -2; done
for mode in 0 1; do for r in 2 3 4 6; do echo "w4 mode$mode R=$r"; COSIM_POOL_MODE=$mode COSIM_POOL_R=$r timeout 300 python tools/soak.py 30 16384 w4_p_v2 stairs_up_hard 2>&1 | tail -1; done; done
for mode in 0 1; do for r in 2 4; do echo "hum mode$mode R=$r"; COSIM_POOL_MODE=$mode COSIM_POOL_R=$r timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1; done; done
