# sky synthesis round: parity of the same-size-source path, timings aligned / general
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "sky" 2>&1 | tail -5
python tools/time_sky.py 2048 2048 2>&1 | tail -2
python tools/time_sky.py 2048 1024 2>&1 | tail -2
python tools/time_sky.py 2048 4096 2>&1 | tail -2
