#!/bin/bash
TAG=${1:-r01e}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench c2"; timeout 300 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_c2.json | cut -c1-240
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_ref.json | cut -c1-200
for wl in c1 c3 c4; do echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_${wl}.json | cut -c1-200; done
echo "== bench c5"; timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --cpu-seconds 6 --ref-batch 32 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_c5.json | cut -c1-200
echo "== ncu launches"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-100
echo "== ncu full (wave c2)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ctc_wave" -s 95 -c 1 -o gpurun_out/prof_wave_c2_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log
ncu -i gpurun_out/prof_wave_c2_${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_wave_c2_${TAG}_source.csv 2>/dev/null
python scratch/ncu_lines.py gpurun_out/prof_wave_c2_${TAG}_source.csv 60 > gpurun_out/ncu_wave_c2_${TAG}_lines.txt 2>&1
