#!/bin/bash
# time kernel variants: one bench line per variant (kernel_ms = lookup+main+finish over 2^20 resident records)
for f in tools/variants/libvar_*.so hotstuff_b200/libhs_crypto.so; do
  HS_CRYPTO_LIB=$PWD/$f python bench.py --steps 3 --warmup 3 --key-mode indexed --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print('$f', 'step_ms=%.3f kernel_ms=%.3f value=%.3e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']))"
done
