for look in 2 32 4 2 32 4; do
  for K in 4 1; do
  MRC_B200_SPARSE_LOOK=$look python bench.py --sparse --no-cpu-baseline --klocal $K --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('look $look K $K', round(d['ms_per_step'],3), 'ms rounds', d['rounds'], 'e2e ms', round(d['e2e']['ms_per_step'],2))"
  done
done
