#!/bin/bash
# usage: scratch/gpu_round.sh <tag>   (runs on the GPU box under gpurun)
TAG=${1:-r01}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c2"; timeout 300 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_c2.json
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_ref.json
for wl in c1 c3 c4; do echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_${wl}.json; done
echo "== bench c5"; timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --cpu-seconds 6 --ref-batch 32 2>&1 | tail -3 | tee gpurun_out/bench_${TAG}_c5.json
echo "== ncu launches"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${TAG}.log
echo "== ncu full (lattice c2)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ctc_lattice|ctc_sweep" -s 5 -c 2 -o gpurun_out/prof_lattice_c2_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
echo "== ncu full (lattice c4)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ctc_lattice|ctc_sweep" -s 5 -c 1 -o gpurun_out/prof_lattice_c4_${TAG} -f python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "== ncu full (grad, rowstats c4)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ctc_grad|ctc_row_stats" -s 10 -c 2 -o gpurun_out/prof_dense_c4_${TAG} -f python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -20
