set -x
mkdir -p gpurun_out
T=r1v9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload cfg4 --steps 5 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg4.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload cfg5 --steps 5 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg5.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full.log 2>&1
python -c "
import json
for f in ('cfg3','cfg4','cfg5'):
    d=json.load(open('gpurun_out/r1v9_bench_%s.json'%f)); print(f,'ms/step %.4f value %.3e e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), round(d['roofline']['frac'],3))"
