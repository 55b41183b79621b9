#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/microbench/fe_warp 2>&1 | tee gpurun_out/r2_fe_warp.txt
python bench.py --key-mode generic --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/r2_bench_generic.json 2>/dev/null; python tools/variants/lastjson.py gpurun_out/r2_bench_generic.json
python bench.py --key-mode cache --steps 5 --warmup 4 --no-cpu-baseline --no-strong > gpurun_out/r2_bench_cache.json 2>/dev/null; python tools/variants/lastjson.py gpurun_out/r2_bench_cache.json
ncu --set full --clock-control none --import-source on -k regex:"k_verify_main" -s 1 -c 1 -o gpurun_out/r2_prof_generic python tools/prof_run.py 131072 generic > gpurun_out/r2_ncu_generic.log 2>&1
ncu -i gpurun_out/r2_prof_generic.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for m in ('gpu__time_duration.sum','launch__registers_per_thread','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'):
    if m in h: print(m, [r[h.index(m)] for r in rows[2:]])
" | tee gpurun_out/r2_ncu_generic_summary.txt
