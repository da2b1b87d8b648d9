# experiment: work-item shapes of the dynamic plan (rows per big item, rows per tail item, tail percent) at cfg3
LIBV=${1:-}
for d in "38 14 25" "46 14 25" "54 14 25" "62 14 25" "46 22 25" "46 14 15" "54 14 15" "54 22 35" "78 14 25"; do set -- $d
echo -n "dyn $1 small $2 tail $3: "; WT_LIB=$LIBV WT_DYN=$1 WT_DYN_SMALL=$2 WT_DYN_TAIL=$3 python tools/time_variants.py 4096:64:shade,fast | grep -v "^$"; done
