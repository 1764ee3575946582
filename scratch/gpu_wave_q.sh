#!/bin/bash
export OMP_NUM_THREADS=16
echo "== check"; E2E_CTC_WAVE=1 timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | cut -c1-110
echo "== default"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-80
echo "== NC=6"; E2E_CTC_WAVE_NC=6 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
echo "== NC=3"; E2E_CTC_WAVE_NC=3 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
