set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_synth -s 4 -c 1 -f -o gpurun_out/r2_prof_sky python tools/time_sky.py 2048 2048 > gpurun_out/ncu_sky.log 2>&1
tail -3 gpurun_out/ncu_sky.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "sky" 2>&1 | tail -3
