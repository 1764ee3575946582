#!/bin/bash
export OMP_NUM_THREADS=16
export E2E_CTC_WAVE=1
echo "== check"; timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | cut -c1-110
echo "== default (NC=4 NP=2)"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-80
echo "== NP=1"; E2E_CTC_WAVE_NP=1 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
echo "== NP=4"; E2E_CTC_WAVE_NP=4 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
echo "== NC=2"; E2E_CTC_WAVE_NC=2 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
echo "== NC=6"; E2E_CTC_WAVE_NC=6 timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "per-call" | cut -c1-80
