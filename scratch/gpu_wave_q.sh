#!/bin/bash
export OMP_NUM_THREADS=16
for cfg in "R=128" "R=256" "R=256 NC=5" "R=256 NC=4" "R=256 NP=4"; do
  envs=""; for kv in $cfg; do envs="$envs E2E_CTC_WAVE_${kv}"; done
  echo "== $cfg"; env $envs timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-80
done
