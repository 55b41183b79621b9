#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest2.log; cat gpurun_out/r2_pytest2.log
./tools/replay_config5 1000 > gpurun_out/r2_replay_config5.json 2> gpurun_out/r2_replay_config5.err; cat gpurun_out/r2_replay_config5.json; tail -3 gpurun_out/r2_replay_config5.err
bash tools/variants/run.sh indexed 2>&1 | tee gpurun_out/r2_variants.txt
ncu --set full --clock-control none --import-source on -k regex:"k_verify_main|k_digest32|k_verify_finish|k_key_lookup" -s 5 -c 5 -o gpurun_out/r2_prof python tools/prof_run.py 1048576 committee > gpurun_out/r2_ncu.log 2>&1
python tools/ncu_traffic.py gpurun_out/r2_prof.ncu-rep 1048576 gpurun_out/r2_ncu_summary.md gpurun_out/r2_traffic.json
