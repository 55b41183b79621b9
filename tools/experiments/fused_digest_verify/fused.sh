#!/bin/bash
# Digest + verify in one launch (k_digest_verify) against the two-kernel form, same box: parity test, then bench lines per lead.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch" > gpurun_out/r2_fused_pytest.log 2>&1; tail -3 gpurun_out/r2_fused_pytest.log
for cfg in "0 1024" "1 1024" "1 592" "1 2048" "1 256" "1 4096"; do
  set -- $cfg
  HS_FUSED=$1 HS_FUSED_LEAD=$2 timeout 200 python bench.py --no-strong --no-cpu-baseline > gpurun_out/r2_fused_$1_$2.json 2>gpurun_out/r2_fused_$1_$2.err
  echo "fused=$1 lead=$2: $(python tools/variants/lastjson.py gpurun_out/r2_fused_$1_$2.json)"
done
