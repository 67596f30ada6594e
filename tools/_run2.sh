python tools/phase_profile.py w4_p_v2 stairs_up_hard 4096 3 > gpurun_out/r2b_phase_w4.log 2>&1; cat gpurun_out/r2b_phase_w4.log
python tools/phase_profile.py flamingo_p_v3 rocky_hard 16384 10 > gpurun_out/r2b_phase_fl.log 2>&1; cat gpurun_out/r2b_phase_fl.log
python tools/phase_profile.py humanoid_p_v0 slope_hard 8192 5 > gpurun_out/r2b_phase_hum.log 2>&1; cat gpurun_out/r2b_phase_hum.log
