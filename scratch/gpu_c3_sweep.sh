#!/bin/bash
export OMP_NUM_THREADS=16
run() { timeout 300 python bench.py --workload ${WL:-c3} --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/b.json; python - <<PY
import json
try:
  d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
  print('$1', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'utt/s %.0f'%d['value'])
except Exception as e: print('$1 failed', open('/tmp/b.json').read()[-300:])
PY
}
E2E_CTC_WAVE=0 run "sweep"
E2E_CTC_WAVE=1 run "wave default (NC2 NP1)"
E2E_CTC_WAVE=1 E2E_CTC_WAVE_NP=2 run "wave NP2"
E2E_CTC_WAVE=1 E2E_CTC_WAVE_NP=4 run "wave NP4"
E2E_CTC_WAVE=1 E2E_CTC_WAVE_NP=2 E2E_CTC_WAVE_NC=1 run "wave NP2 NC1"
E2E_CTC_WAVE=1 E2E_CTC_WAVE_NP=2 E2E_CTC_WAVE_NC=3 run "wave NP2 NC3"
E2E_CTC_WAVE=1 E2E_CTC_WAVE_NP=2 E2E_CTC_WAVE_RV=16 run "wave NP2 RV16"
