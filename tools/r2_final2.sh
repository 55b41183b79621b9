#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err
python tools/variants/lastjson.py gpurun_out/r2_bench_final_1gpu.json
python tools/h2d_bw.py
ncu --set full --clock-control none --import-source on -k regex:"k_verify_main|k_verify_finish|k_digest32_fixed|k_key_lookup" -s 5 -c 5 -o gpurun_out/r2_prof_final python tools/prof_run.py 1048576 committee > gpurun_out/r2_ncu_final.log 2>&1
python tools/ncu_traffic.py gpurun_out/r2_prof_final.ncu-rep 1048576 gpurun_out/r2_ncu_summary_final.md gpurun_out/r2_traffic_final.json | cut -c1-400
