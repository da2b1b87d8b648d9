# round 2, first GPU call: full GPU suite (all failures shown), smoke, margin table, the bench line, variant timings
set -x
mkdir -p gpurun_out
T=${TAG:-r2a}
timeout 2400 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 1200 python tools/parity_margins.py gpurun_out/${T}_parity_margins.json > gpurun_out/${T}_margins.log 2>&1; tail -3 gpurun_out/${T}_margins.log | cut -c1-200
timeout 900 python bench.py --steps 20 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json
timeout 300 python bench.py --steps 50 --phase recurrence --no-extra --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_recur.json 2>> gpurun_out/${T}_bench.err
{
echo "== production"; python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade 4096:64:shade,fast 8192:256: 16384:32: 1024:8:fast
echo "== hmpacked"; WT_LIB=wave_tool_b200/variants/libwt_hmpacked.so python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade 4096:64:shade,fast
} > gpurun_out/${T}_variants.log 2>&1
cat gpurun_out/${T}_variants.log
