# round 2, GPU call B: the suite after the three-tier AUTO change, margins incl. the reference-order kernel, variant parity + timings
set -x
mkdir -p gpurun_out
T=${TAG:-r2b}
timeout 2400 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-120
timeout 1500 python tools/parity_margins.py gpurun_out/${T}_parity_margins.json --quick > gpurun_out/${T}_margins.log 2>&1; tail -2 gpurun_out/${T}_margins.log | cut -c1-200
for n in seamless pad sp; do
  WT_LIB=wave_tool_b200/variants/libwt_$n.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "cfg1 or cfg2 or cfg3_shaded or reference_order_kernel or ragged or wave_count or row_sharded or varyings or sky_footprint or no_transcendentals or large_grid_digest" 2>&1 | tail -4 > gpurun_out/${T}_${n}_pytest.log; tail -2 gpurun_out/${T}_${n}_pytest.log
done
{
echo "== production"; python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade 4096:64:shade,fast 4096:64:shade,ref 8192:256: 8192:256:fast 16384:32: 1024:8:fast 1024:8:
for n in seamless pad sp bands1; do
echo "== $n"; WT_LIB=wave_tool_b200/variants/libwt_$n.so python tools/time_variants.py 4096:64:shade 4096:64:shade,fast 4096:64:shade 4096:64:shade,fast 8192:256: 8192:256:fast
done
} > gpurun_out/${T}_variants.log 2>&1
cat gpurun_out/${T}_variants.log | grep -v "^+"
