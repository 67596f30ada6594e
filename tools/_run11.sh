python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
COSIM_SELFCOL=0 python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
timeout 300 python tools/soak.py 60 16384 w4_p_v2 stairs_up_hard 2>&1 | tail -1
