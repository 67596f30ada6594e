python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
COSIM_SELFCOL=0 python tools/quick_rate.py 65536 20 5 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_pytest.log 2>&1; tail -40 gpurun_out/r2i_pytest.log
