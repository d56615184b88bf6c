for m in none copy compute d2h; do
  if [ $m = none ]; then f=""; else f="--e2e-ablate $m"; fi
  timeout 200 python bench.py --no-cpu-baseline --no-parity --batch8 0 $f 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$m', 'value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
done
