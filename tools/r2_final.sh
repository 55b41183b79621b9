#!/bin/bash
# final single-GPU evidence of round 2: tests, default bench + reference arm, replay, launch list, sanitizer
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_pytest_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference_arm.json 2>/dev/null
python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err
python tools/variants/lastjson.py gpurun_out/r2_bench_final_1gpu.json gpurun_out/r2_bench_final_reference_arm.json
./tools/replay_config5 1000 > gpurun_out/r2_replay_config5.json 2>/dev/null; cat gpurun_out/r2_replay_config5.json | cut -c1-900
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
