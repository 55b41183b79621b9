#!/bin/bash
# time kernel variants: one bench line per variant (kernel_ms = lookup+main+finish over 2^20 resident records)
MODE=${1:-indexed}
for f in tools/variants/libvar_*.so hotstuff_b200/libhs_crypto.so; do
  HS_CRYPTO_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --key-mode $MODE --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/var.json
  python tools/variants/lastjson.py /tmp/var.json | sed "s|/tmp/var.json|$f|"
done
