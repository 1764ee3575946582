#!/bin/bash
# usage: scratch/gpu_round.sh <tag>   (runs on the GPU box under gpurun)
TAG=${1:-r01}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c2"; timeout 300 python bench.py --steps 200 --warmup 10 2>&1 | tail -2 | tee gpurun_out/bench_${TAG}_c2.json
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_ref.json
for wl in c1 c3 c4; do echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_${wl}.json; done
echo "== ncu launches"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_launch_${TAG}.log
echo "== ncu full (lattice)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ctc_lattice -s 5 -c 2 -o gpurun_out/prof_lattice_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ctc_grad -s 5 -c 1 -o gpurun_out/prof_grad_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -20
