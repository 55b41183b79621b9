#!/bin/bash
# k_digest_verify with atomicAdd claims: parity, bench with / without it on the same box, then the whole GPU suite with it on
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch" > gpurun_out/r2_fuseda_pytest.log 2>&1; tail -2 gpurun_out/r2_fuseda_pytest.log
HS_FUSED=1 timeout 100 python bench.py --no-strong --no-cpu-baseline > gpurun_out/r2_fuseda_on.json 2>gpurun_out/r2_fuseda_on.err
echo "on: $(python tools/variants/lastjson.py gpurun_out/r2_fuseda_on.json)"
HS_FUSED=0 timeout 100 python bench.py --no-strong --no-cpu-baseline --no-e2e > gpurun_out/r2_fuseda_off.json 2>gpurun_out/r2_fuseda_off.err
echo "off: $(python tools/variants/lastjson.py gpurun_out/r2_fuseda_off.json)"
HS_FUSED=1 timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_fuseda_pytest_all.log 2>&1; tail -2 gpurun_out/r2_fuseda_pytest_all.log
