mkdir -p gpurun_out
T=r1v5
timeout 600 python bench.py --exact-phase --no-cpu-baseline > gpurun_out/${T}_bench_cfg3_exact.json 2>> gpurun_out/${T}_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_exact python bench.py --exact-phase --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full2.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "cfg3 or shard or ragged" 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open("gpurun_out/r1v5_bench_cfg3_exact.json"))
print("exact ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]), d["roofline"]["bound"], round(d["roofline"]["frac"],3))
PY
