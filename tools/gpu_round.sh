mkdir -p gpurun_out
T=r1v4
python tools/explore_sched.py --tag base --mode fast --configs=-1:0,14:0,22:0,30:0,38:0,46:0,62:0 2>&1 | grep '"tag"' | cut -c1-120
timeout 600 python bench.py --workload cfg4 --steps 5 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload cfg5 --steps 5 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg5.json 2>> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload cfg2 --steps 200 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2>> gpurun_out/${T}_bench.err
python - <<'PY'
import json
for f in ["cfg4","cfg5","cfg2"]:
    try:
        d=json.load(open(f"gpurun_out/r1v4_bench_{f}.json"))
        print(f, "ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d["roofline"]["bound"], round(d["roofline"]["frac"],3), d["scaling"])
    except Exception as e: print(f, "failed", e)
PY
