#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
export E2E_CTC_WAVE=1
export E2E_CTC_WAVE_BY_SMSP=0
for r in 1; do echo "== wave check $r"; timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | cut -c1-110; done
echo "== per-call"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-110
echo "== pytest"; E2E_CTC_WAVE=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
