for mk in 63 575 1599 2623 3647 3775; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 40 60 2>&1 | tail -1; done
for mk in 63 3775; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 20 5 2>&1 | tail -1; done
for mk in 63 575 3647; do echo "mask $mk"; COSIM_SOAK_BLOCK=30 COSIM_BSYNC_MASK=$mk timeout 300 python tools/soak.py 60 65536 flamingo_light_v1 flat 2>&1 | tail -1 | cut -c1-60; done
COSIM_BSYNC_MASK=3775 python tools/pool_check.py 2>&1 | tail -1
