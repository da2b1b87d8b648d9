for k in 20 200 2000; do
timeout 600 python bench.py --steps $k --no-cpu-baseline --no-readback 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['steps'], 'ms/step %.4f value %.3e e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), d['clocks'])"
done
