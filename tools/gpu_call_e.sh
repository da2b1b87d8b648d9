set -x
mkdir -p gpurun_out
T=${TAG:-r2e}
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-250
python tools/time_sky.py > gpurun_out/${T}_sky.log 2>&1; tail -3 gpurun_out/${T}_sky.log
python tools/time_compose.py > gpurun_out/${T}_compose.log 2>&1; tail -4 gpurun_out/${T}_compose.log
