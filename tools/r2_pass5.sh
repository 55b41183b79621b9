#!/bin/bash
set -u
mkdir -p gpurun_out
HS_KEY_CACHE=0 ncu --set full --clock-control none --import-source on -k regex:"k_verify_main" -c 6 -o gpurun_out/r2_prof_generic python tools/prof_run.py 262144 generic > gpurun_out/r2_ncu_generic.log 2>&1
ncu -i gpurun_out/r2_prof_generic.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for m in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'):
    if m in h: print(m, [r[h.index(m)][:24] for r in rows[2:]])
" | tee gpurun_out/r2_ncu_generic_summary.txt
