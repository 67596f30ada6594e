for lib in _build/libcosim_b200.so _build_ab/lib_nopgs.so _build_ab/lib_head.so _build/libcosim_b200.so _build_ab/lib_head.so; do
  echo $lib; COSIM_LIB_PATH=cosim_b200/csrc/$lib python tools/quick_rate.py 65536 40 60 2>&1 | tail -1
done
