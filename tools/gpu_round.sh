mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6n_pytest.log 2>&1; tail -4 gpurun_out/s6n_pytest.log
timeout 600 python bench.py > gpurun_out/r1v6_bench_cfg3.json 2> gpurun_out/r1v6_bench.err; tail -3 gpurun_out/r1v6_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r1v6_bench_cfg3.json"))
print("ms/step %.4f value %.3e e2e %.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d["roofline"]["bound"], round(d["roofline"]["frac"],3), d["gpu_launches"], d["cpu_baseline"]["value"])
PY
python tools/time_sky.py 2048 2048
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
