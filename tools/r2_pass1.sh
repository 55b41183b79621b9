#!/bin/bash
# round-2 GPU pass 1: parity suite, default bench, SHA variant timing, launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.err
bash tools/variants/run.sh 2>&1 | tee gpurun_out/r2_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -c . gpurun_out/r2_launches.csv
python tools/variants/lastjson.py gpurun_out/r2_bench_default.json
