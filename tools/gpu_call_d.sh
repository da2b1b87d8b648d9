# round 2, GPU call D: sky synthesis (new approximate path) tests + timing, work-item shape sweep of the bit-faithful kernel
set -x
mkdir -p gpurun_out
T=${TAG:-r2d}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --tb=short -p no:cacheprovider -k "skybox or heightmap_sample_scale or step_host" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -14 gpurun_out/${T}_pytest.log | cut -c1-200
python tools/time_sky.py > gpurun_out/${T}_sky.log 2>&1; cat gpurun_out/${T}_sky.log | tail -6
python tools/time_compose.py > gpurun_out/${T}_compose.log 2>&1; tail -4 gpurun_out/${T}_compose.log
V=wave_tool_b200/variants/libwt_dev.so
{
for d in "30 0 0" "46 14 15" "62 14 15" "62 14 25" "46 22 20" "94 14 20" "62 30 25" "30 0 0"; do set -- $d
echo -n "exact dyn $1 small $2 tail $3: "; WT_LIB=$V WT_DYN=$1 WT_DYN_SMALL=$2 WT_DYN_TAIL=$3 python tools/time_variants.py 4096:64:shade | grep -v "^$"; done
} > gpurun_out/${T}_sweep_exact.log 2>&1
grep -v "^+" gpurun_out/${T}_sweep_exact.log
