timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; tail -6 gpurun_out/r2q_pytest.log
for lib in _build/libcosim_b200.so _build_ab/lib_head.so _build/libcosim_b200.so; do
  echo $lib; COSIM_LIB_PATH=cosim_b200/csrc/$lib python tools/quick_rate.py 65536 40 60 2>&1 | tail -1
done
