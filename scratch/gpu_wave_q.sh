#!/bin/bash
export OMP_NUM_THREADS=16
echo "== check"; E2E_CTC_WAVE=1 timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | cut -c1-110
echo "== default"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-80
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for ch in 2 3 4 6 8; do echo "host chunks=$ch"; E2E_CTC_HOST_CHUNKS=$ch timeout 100 python scratch/host_overhead3.py c2 2>&1 | head -1; done
