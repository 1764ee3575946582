#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
echo "== wave check"; timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | head
echo "== dbg B=1"; E2E_CTC_WAVE_DBG=1 timeout 200 python scratch/gpu_wave_dbg2.py 1 2>&1 | tail -12
echo "== per-call"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-110
