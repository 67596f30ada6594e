for mk in 63 447 575 703; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 40 60 2>&1 | tail -1; done
for mk in 447 703; do echo "mask $mk"; COSIM_SOAK_BLOCK=30 COSIM_BSYNC_MASK=$mk timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1 | cut -c1-60; done
COSIM_BSYNC_MASK=703 python tools/pool_check.py 2>&1 | tail -1
