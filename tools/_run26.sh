for mk in 63 319; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 40 60 2>&1 | tail -1; done
for mk in 63 319; do echo "mask $mk"; COSIM_BSYNC_MASK=$mk python tools/quick_rate.py 65536 20 5 2>&1 | tail -1; done
