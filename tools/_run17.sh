python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
timeout 300 python tools/soak.py 100 65536 2>&1 | tail -1
python tools/pool_check.py 8192 10 2>&1 | tail -1
ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,sass__inst_executed_local_stores,sass__inst_executed_local_loads --clock-control none -k regex:k_step -s 3 -c 1 --csv --log-file gpurun_out/r2n_dram.csv python tools/prof_run.py 16384 4 > /dev/null 2>&1; cat gpurun_out/r2n_dram.csv | tail -6
