python tools/diag_general.py cone=elliptic 2>&1 | tail -8
python tools/quick_rate.py 65536 40 60 2>&1 | tail -1
python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
COSIM_SOAK_BLOCK=30 timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1 | cut -c1-60
COSIM_SOAK_BLOCK=30 timeout 300 python tools/soak.py 60 65536 flamingo_light_v1 flat 2>&1 | tail -1 | cut -c1-60
