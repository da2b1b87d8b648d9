set -x
mkdir -p gpurun_out
T=${TAG:-r2n8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; tail -2 gpurun_out/${T}_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_full_readback',{}).get('value'), json.dumps(d['extra_workloads'])[:1800])"
