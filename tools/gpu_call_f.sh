set -x
mkdir -p gpurun_out
T=${TAG:-r2f}
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log | cut -c1-250
python tools/time_sky.py > gpurun_out/${T}_sky.log 2>&1; tail -3 gpurun_out/${T}_sky.log
python tools/time_variants.py 4096:64:shade,stencil 4096:64:shade,ref 4096:64:shade 1024:8:stencil 8192:256:stencil > gpurun_out/${T}_variants.log 2>&1; grep -v "^+" gpurun_out/${T}_variants.log
