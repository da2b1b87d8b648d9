mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_ -s 6 -c 2 -f -o gpurun_out/s6_sky python tools/time_sky.py 2048 2048 > gpurun_out/s6_sky.log 2>&1
tail -3 gpurun_out/s6_sky.log
