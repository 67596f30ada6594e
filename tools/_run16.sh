sh tools/build_prof.sh 2>&1 | grep -i error
python tools/phase_profile.py w4_p_v2 stairs_up_hard 16384 3 2>&1 | grep -v histogram > gpurun_out/r2m_phase_w4_pool.log; cat gpurun_out/r2m_phase_w4_pool.log
python tools/soak.py 3 8192 w4_p_v2 stairs_up_hard && ncu --set full --clock-control none --import-source on -k regex:k_step -s 2 -c 1 -o gpurun_out/r2m_w4 python tools/soak.py 3 8192 w4_p_v2 stairs_up_hard > gpurun_out/r2m_ncu.log 2>&1; tail -2 gpurun_out/r2m_ncu.log
cp cosim_b200/csrc/_build/engine.cu.o gpurun_out/r2m_engine.cu.o
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
