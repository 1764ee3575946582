#!/bin/bash
# quick iteration: parity tests + short benches (no CPU baseline)
TAG=${1:-q}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for wl in c2 c1 c3 c4 c5; do
  st=100; [ $wl = c5 ] && st=5
  timeout 300 python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${TAG}_${wl}.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_${wl}.json').read().strip().splitlines()[-1])
print('$wl', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'utt/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'])
PY
done
