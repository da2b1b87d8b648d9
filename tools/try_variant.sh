# experiment runner: parity suite + timings of a tuning variant against the production library
#   gpurun -- 'bash tools/try_variant.sh NAME [NAME2 ...]'   (variants built by tools/build_variant.py)
for n in "$@"; do
V=wave_tool_b200/variants/libwt_$n.so
WT_LIB=$V timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${n}_pytest.log; tail -4 gpurun_out/${n}_pytest.log
done
{
echo "== production"; python tools/time_variants.py 4096:64:shade,fast 4096:64:shade,fast 8192:256:fast 16384:32:fast 1024:8:fast
for n in "$@"; do
V=wave_tool_b200/variants/libwt_$n.so
echo "== $n"; WT_LIB=$V python tools/time_variants.py 4096:64:shade,fast 4096:64:shade,fast 4096:64:shade,fast,cks 8192:256:fast 16384:32:fast 1024:8:fast
echo "== $n wave loop only"; WT_DBG=1 WT_LIB=$V python tools/time_variants.py 4096:64:shade,fast
done
} 2>&1 | grep -v "^$" | tee gpurun_out/variant_timing.log
