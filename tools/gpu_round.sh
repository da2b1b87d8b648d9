mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6k_pytest.log 2>&1; tail -25 gpurun_out/s6k_pytest.log
python tools/explore_sched.py --tag base --mode fast,exact --configs -1:0 2>&1 | tail -3
