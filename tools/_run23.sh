timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; tail -4 gpurun_out/r2t_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
