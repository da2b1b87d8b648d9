set -x
mkdir -p gpurun_out
T=${TAG:-r2h}
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log | cut -c1-250
RING=4 python tools/time_animation.py 1024 8 1000 > gpurun_out/${T}_anim4.log 2>&1; grep -v "^+" gpurun_out/${T}_anim4.log | tail -3
RING=8 python tools/time_animation.py 1024 8 1000 > gpurun_out/${T}_anim8.log 2>&1; grep -v "^+" gpurun_out/${T}_anim8.log | tail -3
python tools/time_variants.py 4096:64:shade,stencil 1024:8:fast 1024:8: 4096:64:shade > gpurun_out/${T}_variants.log 2>&1; grep -v "^+" gpurun_out/${T}_variants.log
timeout 900 python bench.py --steps 20 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], json.dumps(d['extra_workloads'])[:1200])"
