N=${1:-2}
for coll in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --collective $coll --no-e2e --no-cpu-baseline > gpurun_out/mg_msgs_${N}_$coll.json 2>gpurun_out/mg_msgs_${N}_$coll.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload qc --committee 10000 --qcs 150 --votes-per-qc 6667 --steps 10 --warmup 3 --collective $coll > gpurun_out/mg_qc_${N}_$coll.json 2>gpurun_out/mg_qc_${N}_$coll.err
done
python tools/variants/lastjson.py gpurun_out/mg_*_${N}_*.json
