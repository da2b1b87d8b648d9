mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "step_host" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/final_bench.json 2>gpurun_out/final_bench.err; tail -3 gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('ms/step %.4f value %.3e e2e %.3e sync %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['synchronous_value']))"
