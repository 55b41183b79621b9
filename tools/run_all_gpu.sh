#!/bin/bash
# Reproduces the round's GPU evidence on a 1-GPU B200 box (run from the repo root, e.g. under gpurun):
#   tests -> default bench (+ reference arm) -> alternative key modes -> QC workload -> ncu launch list + full capture -> summaries.
# Multi-GPU points: bash tools/variants/mg2.sh N   (N = 2, 4, 8).
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for mode in indexed cache generic; do
  python bench.py --key-mode $mode --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$mode.json 2>/dev/null
done
python bench.py --workload qc --steps 5 --warmup 3 > gpurun_out/bench_qc.json 2>/dev/null
python tools/replay_config5.py 1000 > gpurun_out/replay_config5.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_verify_main|k_digest32|k_verify_finish|k_key_lookup" -s 5 -c 5 -o gpurun_out/prof python tools/prof_run.py 262144 committee > gpurun_out/ncu.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof.ncu-rep 262144 gpurun_out/ncu_summary.md gpurun_out/traffic.json
python tools/variants/lastjson.py gpurun_out/bench_default.json gpurun_out/bench_reference.json gpurun_out/bench_indexed.json gpurun_out/bench_cache.json gpurun_out/bench_generic.json gpurun_out/bench_qc.json
