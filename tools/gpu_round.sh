mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1v8_bench_cfg3.json 2>gpurun_out/r1v8_bench.err; tail -2 gpurun_out/r1v8_bench.err
timeout 600 python bench.py --exact-phase --no-cpu-baseline > gpurun_out/r1v8_bench_cfg3_exact.json 2>>gpurun_out/r1v8_bench.err
python -c "
import json
for f in ('cfg3','cfg3_exact'):
    d=json.load(open('gpurun_out/r1v8_bench_%s.json'%f)); print(f,'ms/step %.4f value %.3e e2e %.3e sync %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['synchronous_value']), round(d['roofline']['frac'],3), d['clocks'])"
