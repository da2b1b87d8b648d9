# round 2, GPU call C: suite, nomagic A/B, ncu launch list + full captures of both production kernels, the bench line
set -x
mkdir -p gpurun_out
T=${TAG:-r2c}
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-250
{
echo "== production"; python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade 4096:64:shade,fast 8192:256: 16384:32:
echo "== nomagic"; WT_LIB=wave_tool_b200/variants/libwt_nomagic.so python tools/time_variants.py 4096:64:shade 4096:64:shade 8192:256: 16384:32:
echo "== production again"; python tools/time_variants.py 4096:64:shade 4096:64:shade
} > gpurun_out/${T}_variants.log 2>&1
grep -v "^+" gpurun_out/${T}_variants.log
timeout 900 python bench.py --steps 20 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-readback --no-extra --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg3.csv $B > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_exact $B > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_warp -s 3 -c 2 -f -o gpurun_out/${T}_prof_fast $B --phase recurrence > gpurun_out/${T}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -8
