mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "compose" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6m_pytest.log 2>&1; tail -5 gpurun_out/s6m_pytest.log
