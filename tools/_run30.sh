for L in "" cosim_b200/csrc/_build_ab/lib_fwd.so cosim_b200/csrc/_build_ab/lib_revdense.so; do echo "lib [$L]";
COSIM_LIB_PATH=$L python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
COSIM_SOAK_BLOCK=40 COSIM_LIB_PATH=$L python tools/soak.py 80 65536 2>&1 | tail -1
COSIM_SOAK_BLOCK=30 COSIM_LIB_PATH=$L timeout 300 python tools/soak.py 60 16384 humanoid_p_v0 slope_hard 2>&1 | tail -1
done
