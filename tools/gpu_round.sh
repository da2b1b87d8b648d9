# One GPU-box round: parity suite, smoke, the bench line (both phase modes), launch list and ncu captures.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'        (outputs under gpurun_out/, summaries via tools/ncu_summary.py)
set -x
mkdir -p gpurun_out
T=${TAG:-round}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --exact-phase --no-cpu-baseline > gpurun_out/${T}_bench_cfg3_exact.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_exact python bench.py --exact-phase --steps 3 --warmup 3 --no-cpu-baseline --no-readback > gpurun_out/${T}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
# the other configurations (not the bench line), the reference arm and the sanitizer passes
for c in cfg4 cfg5; do timeout 300 python bench.py --workload $c --steps 10 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_$c.json 2>> gpurun_out/${T}_bench.err; done
timeout 300 python bench.py --workload cfg2 --steps 200 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg2.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wave_count_tiers or varyings" 2>&1 | tail -4 > gpurun_out/${T}_racecheck.log
tail -2 gpurun_out/${T}_memcheck.log gpurun_out/${T}_racecheck.log
