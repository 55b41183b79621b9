#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest3.log; cat gpurun_out/r2_pytest3.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 1500 gpurun_out/r2_bench_default.err; python tools/variants/lastjson.py gpurun_out/r2_bench_default.json
python bench.py --base-window 24 --no-strong --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_w24.json 2>/dev/null; python tools/variants/lastjson.py gpurun_out/r2_bench_w24.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; python tools/variants/lastjson.py gpurun_out/r2_bench_reference.json
