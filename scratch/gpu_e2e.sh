#!/bin/bash
export OMP_NUM_THREADS=16
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for wl in c2 c3 c4 c1; do
  for ch in -1 1; do
  E2E_CTC_HOST_CHUNKS=$ch timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/b.json
  python - <<PY
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('$wl chunks=$ch', 'ms/step %.3f'%d['ms_per_step'], 'utt/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'])
PY
  done
done
