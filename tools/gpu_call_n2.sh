set -x
mkdir -p gpurun_out
T=${TAG:-r2n2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; tail -3 gpurun_out/${T}_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_full_readback'), json.dumps(d['extra_workloads'])[:1500])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2>> gpurun_out/${T}_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/${T}_ref.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --workload cfg5 --no-extra > gpurun_out/${T}_cfg5.json 2>> gpurun_out/${T}_bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/${T}_cfg5.json
