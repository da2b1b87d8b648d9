mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6l_pytest.log 2>&1; tail -5 gpurun_out/s6l_pytest.log
python tools/time_animation.py 1024 8 1000
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s6l_launches_cfg2.csv python tools/time_animation.py 1024 8 20 > /dev/null 2>&1
grep surface_warp gpurun_out/s6l_launches_cfg2.csv | tail -8 | cut -d, -f5,12-
