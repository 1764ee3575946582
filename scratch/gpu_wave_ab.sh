#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
for wl in ${WLS:-c2 c1 c3 c5}; do
  for wave in 1 0; do
  st=100; [ $wl = c5 ] && st=5
  E2E_CTC_WAVE=$wave timeout 300 python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_wave${wave}_${wl}.json
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_wave${wave}_${wl}.json').read().strip().splitlines()[-1])
  print('$wl wave=$wave', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'utt/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'])
except Exception as e:
  print('$wl wave=$wave failed', e, open('gpurun_out/bench_wave${wave}_${wl}.json').read()[-300:])
PY
  done
done
