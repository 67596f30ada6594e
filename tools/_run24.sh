for mk in 63 47 39 55 31; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 40 60 2>&1 | tail -1; done
