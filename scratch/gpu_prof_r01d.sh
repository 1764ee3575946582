#!/bin/bash
TAG=r01d
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
echo "== ncu launches (bench c2)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-200
echo "== ncu full (wave c2)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ctc_wave" -s 95 -c 2 -o gpurun_out/prof_wave_c2_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
ncu -i gpurun_out/prof_wave_c2_${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_wave_c2_${TAG}_source.csv 2>/dev/null
python scratch/ncu_lines.py gpurun_out/prof_wave_c2_${TAG}_source.csv 60 > gpurun_out/ncu_wave_c2_${TAG}_lines.txt 2>&1
echo "== A/B c3 c5 c1 wave forced"
WLS="c3 c1" bash scratch/gpu_wave_ab.sh 2>&1 | grep -E "^c[0-9]"
