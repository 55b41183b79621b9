#!/bin/bash
# multi-GPU pass: N = $1 ; runs the 2-GPU peer test (when N >= 2), the default bench (weak headline + strong config[3] leg) with both collectives
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N | tr '\n' ';'; echo
if [ "$N" -ge 2 ] && [ "${2:-}" != "notest" ]; then timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_mg_$N.log; fi
for coll in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 --collective $coll > gpurun_out/r2_scale_${N}_$coll.json 2> gpurun_out/r2_scale_${N}_$coll.err
  tail -c 400 gpurun_out/r2_scale_${N}_$coll.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_scale_${N}_$coll.json").read().splitlines() if l.startswith("{")][-1])
    s=d.get("strong_scaling_config3") or {}
    print("N=$N $coll weak value=%.4e ms=%.3f e2e=%.3e | strong config3 ms=%.3f votes/s=%.3e launches/step=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], s.get("ms_per_step",0), s.get("votes_per_s",0), s.get("gpu_launches_per_step")))
except Exception as e:
    print("N=$N $coll FAILED", e)
PY
done
