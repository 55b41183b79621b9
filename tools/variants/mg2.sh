set -x
for coll in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --collective $coll --no-e2e --no-cpu-baseline 2>gpurun_out/mg2_$coll.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print('msgs $coll', d['value'], d['ms_per_step'], d['config']['collective'])" || tail -5 gpurun_out/mg2_$coll.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload qc --committee 1000 --qcs 10000 --votes-per-qc 100 --steps 10 --warmup 3 --collective $coll 2>gpurun_out/mg2qc_$coll.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print('qc $coll', d['value'], d['ms_per_step'], d['config']['collective'])" || tail -5 gpurun_out/mg2qc_$coll.err
done
