for ch in 32768 65536 131072 262144; do HS_CHUNK_RECORDS=$ch python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print('chunk $ch', 'e2e=%.3e value=%.3e' % (d['e2e']['value'], d['value']))"; done
