#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
export E2E_CTC_WAVE=1
export E2E_CTC_WAVE_BY_SMSP=${BYSMSP:-0}
for r in 1 2; do echo "== wave check $r"; timeout 200 python scratch/gpu_wave_check.py 2>&1 | grep -E "viol [1-9]|nan_eq False|ALL OK|FAILED|Error|error" | cut -c1-110; done
echo "== per-call"; timeout 200 python scratch/gpu_wave_dbg.py 2>&1 | grep -E "nbad|per-call" | cut -c1-110
echo "== ncu full wave c2"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"ctc_wave" -s 4 -c 1 -o gpurun_out/prof_wave_c2 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wave.log 2>&1
tail -2 gpurun_out/ncu_wave.log
ncu -i gpurun_out/prof_wave_c2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_wave_c2_source.csv 2>/dev/null
python scratch/ncu_lines.py gpurun_out/prof_wave_c2_source.csv 70 > gpurun_out/prof_wave_c2_lines.txt 2>&1
head -75 gpurun_out/prof_wave_c2_lines.txt | cut -c1-230
