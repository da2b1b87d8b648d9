set -x
mkdir -p gpurun_out
T=r1v3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err; tail -c 400 gpurun_out/${T}_bench_cfg3.json
timeout 600 python bench.py --exact-phase --no-cpu-baseline > gpurun_out/${T}_bench_cfg3_exact.json 2>> gpurun_out/${T}_bench_cfg3.err
timeout 600 python bench.py --workload cfg2 --steps 200 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2>> gpurun_out/${T}_bench_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_exact python bench.py --exact-phase --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
