# experiment: symmetric-pair wave loop (WT_SYM=1, R=9) against the production library
V=wave_tool_b200/variants/libwt_sym9.so
WT_LIB=$V timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/sym_pytest.log; tail -5 gpurun_out/sym_pytest.log
{
echo "== production"; python tools/time_variants.py 4096:64:shade,fast 4096:64:shade,fast 8192:256:fast 16384:32:fast 1024:8:fast
echo "== sym9"; WT_LIB=$V python tools/time_variants.py 4096:64:shade,fast 4096:64:shade,fast 8192:256:fast 16384:32:fast 1024:8:fast
echo "== sym9 wave loop only"; WT_DBG=1 WT_LIB=$V python tools/time_variants.py 4096:64:shade,fast
echo "== production wave loop only"; WT_DBG=1 python tools/time_variants.py 4096:64:shade,fast
for d in "34 16" "52 16" "43 7" "43 25" "61 16"; do set -- $d; echo "== sym9 dyn $1 small $2"; WT_DYN=$1 WT_DYN_SMALL=$2 WT_LIB=$V python tools/time_variants.py 4096:64:shade,fast; done
} 2>&1 | grep -v "^$" | tee gpurun_out/sym_timing.log
