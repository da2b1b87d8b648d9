# One GPU-box round: parity suite, smoke, margin table, the bench line (AUTO and both forced production kernels), launch list
# and ncu captures, the other workloads, the reference arm and the sanitizer passes.
#   gpurun --timeout 3000 -- 'bash tools/gpu_round.sh'        (outputs under gpurun_out/, summaries via tools/ncu_summary.py)
set -x
mkdir -p gpurun_out
T=${TAG:-round}
timeout 1200 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
timeout 1500 python tools/parity_margins.py gpurun_out/${T}_parity_margins.json > gpurun_out/${T}_margins.log 2>&1; tail -1 gpurun_out/${T}_margins.log
timeout 900 python bench.py --steps 20 > gpurun_out/${T}_bench_cfg3_steps20.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
timeout 900 python bench.py > gpurun_out/${T}_bench_cfg3.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --phase recurrence --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_recurrence.json 2>> gpurun_out/${T}_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback --no-extra --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg3.csv $B > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_exact $B > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_fast $B --phase recurrence > gpurun_out/${T}_ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_synth -s 4 -c 1 -f -o gpurun_out/${T}_prof_sky python tools/time_sky.py 2048 2048 > gpurun_out/${T}_ncu_sky.log 2>&1
ls -la gpurun_out | tail -12
# the other configurations (not the bench line), the reference arm and the sanitizer passes
for c in cfg4 cfg5; do timeout 300 python bench.py --workload $c --steps 10 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_$c.json 2>> gpurun_out/${T}_bench.err; done
timeout 300 python bench.py --workload cfg2 --steps 200 --no-cpu-baseline --no-readback > gpurun_out/${T}_bench_cfg2.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
python tools/time_sky.py 2048 2048 > gpurun_out/${T}_sky.log 2>&1; python tools/time_sky.py 2048 1024 >> gpurun_out/${T}_sky.log 2>&1; python tools/time_compose.py > gpurun_out/${T}_compose.log 2>&1; RING=16 python tools/time_animation.py 1024 8 1000 > gpurun_out/${T}_anim.log 2>&1
python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade,stencil 4096:64:shade,ref 8192:256: 8192:256:fast 16384:32: 16384:32:fast 1024:8:fast 1024:8: > gpurun_out/${T}_kernels.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "not cfg3_full_grid and not large_grid and not maximum_grid and not cfg5" 2>&1 | tail -4 > gpurun_out/${T}_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "wave_count_tiers or varyings or cfg2_animation or stream_ordering" 2>&1 | tail -4 > gpurun_out/${T}_racecheck.log
tail -2 gpurun_out/${T}_memcheck.log; tail -2 gpurun_out/${T}_racecheck.log
