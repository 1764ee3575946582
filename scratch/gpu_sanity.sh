#!/bin/bash
export OMP_NUM_THREADS=16
echo "== memcheck (wave check, small cases)"
E2E_CTC_WAVE=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scratch/gpu_wave_check.py tiny c1 nw2 v96 2>&1 | grep -E "ERROR SUMMARY|Invalid|ALL OK|FAILED|out of bounds|misaligned" | head -10
echo "== B=128 and B=148 c2-shaped: wave vs sweep"
for bsz in 96 128 148 256; do for wv in 1 0; do
E2E_CTC_WAVE=$wv timeout 300 python bench.py --workload c2 --batch $bsz --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/b.json
python - <<PY
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('B=$bsz wave=$wv', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, 'utt/s %.0f'%d['value'])
PY
done; done
