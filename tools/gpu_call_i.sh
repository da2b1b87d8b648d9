set -x
mkdir -p gpurun_out
T=${TAG:-r2i}
{
echo "== production"; python tools/time_variants.py 4096:64:shade 4096:64:shade
for n in magic2 lb20 epi2; do echo "== $n"; WT_LIB=wave_tool_b200/variants/libwt_$n.so python tools/time_variants.py 4096:64:shade 4096:64:shade; done
V=wave_tool_b200/variants/libwt_dev.so
for d in "30 0 0" "34 0 0" "38 0 0" "42 0 0" "26 0 0" "38 14 10"; do set -- $d
echo -n "exact dyn $1 small $2 tail $3: "; WT_LIB=$V WT_DYN=$1 WT_DYN_SMALL=$2 WT_DYN_TAIL=$3 python tools/time_variants.py 4096:64:shade | grep -v "^$"; done
} > gpurun_out/${T}_exact_variants.log 2>&1
grep -v "^+" gpurun_out/${T}_exact_variants.log
