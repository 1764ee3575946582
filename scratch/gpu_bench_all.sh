#!/bin/bash
TAG=${1:-r01d}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench c2"; timeout 300 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_c2.json | cut -c1-400
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_ref.json | cut -c1-200
for wl in c1 c3 c4; do echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_${wl}.json | cut -c1-300; done
