#!/bin/bash
# second pass on k_digest_verify: compact SHA body variant, and where the stall cycles of the fused launch go (ncu)
mkdir -p gpurun_out
HS_CRYPTO_LIB=$PWD/tools/variants/libvar_fused_compact.so HS_FUSED=1 timeout 200 python bench.py --no-strong --no-cpu-baseline --no-e2e > gpurun_out/r2_fusedc.json 2>gpurun_out/r2_fusedc.err
echo "compact: $(python tools/variants/lastjson.py gpurun_out/r2_fusedc.json)"
for v in plain compact; do
  lib=$PWD/hotstuff_b200/libhs_crypto.so; [ $v = compact ] && lib=$PWD/tools/variants/libvar_fused_compact.so
  HS_CRYPTO_LIB=$lib HS_FUSED=1 timeout 300 ncu --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section SchedulerStats --section InstructionStats --section Occupancy \
    --clock-control none -k regex:k_digest_verify -s 2 -c 1 -o gpurun_out/r2_fused_$v -f python tools/prof_run.py 1048576 committee > gpurun_out/r2_fused_ncu_$v.log 2>&1
  tail -2 gpurun_out/r2_fused_ncu_$v.log
done
