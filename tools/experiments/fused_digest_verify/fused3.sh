#!/bin/bash
# warp-level k_digest_verify: parity test, bench lines per lead (in 32-record strips), stall picture
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch" > gpurun_out/r2_fusedw_pytest.log 2>&1; tail -3 gpurun_out/r2_fusedw_pytest.log
for lead in 4096 2400 16384; do
  HS_FUSED=1 HS_FUSED_LEAD=$lead timeout 200 python bench.py --no-strong --no-cpu-baseline --no-e2e > gpurun_out/r2_fusedw_$lead.json 2>gpurun_out/r2_fusedw_$lead.err
  echo "warp-level lead=$lead: $(python tools/variants/lastjson.py gpurun_out/r2_fusedw_$lead.json)"
done
HS_FUSED=1 timeout 300 ncu --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section SchedulerStats --section SourceCounters --import-source on \
  --clock-control none -k regex:k_digest_verify -s 2 -c 1 -o gpurun_out/r2_fusedw -f python tools/prof_run.py 1048576 committee > gpurun_out/r2_fusedw_ncu.log 2>&1
tail -2 gpurun_out/r2_fusedw_ncu.log
