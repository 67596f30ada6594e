cd r01_tmp
python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
python tools/phase_profile.py flamingo_p_v3 rocky_hard 16384 10 2>&1 | grep -v histogram
python tools/prof_run.py 16384 4 && ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o ../gpurun_out/r2g_r01 python tools/prof_run.py 16384 4 > ../gpurun_out/r2g_ncu.log 2>&1; tail -2 ../gpurun_out/r2g_ncu.log
