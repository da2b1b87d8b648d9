mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --exact-phase --no-cpu-baseline > gpurun_out/r1v9_bench_cfg3_exact.json 2>gpurun_out/r1v9_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/r1v9_prof_exact python bench.py --exact-phase --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/r1v9_ncu_full2.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r1v9_bench_cfg3_exact.json')); print('exact ms/step %.4f value %.3e e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), round(d['roofline']['frac'],3))"
