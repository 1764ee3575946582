#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
echo "== ubench"; timeout 120 profiles/microbench/ubench 2>&1 | tee gpurun_out/ubench.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for wl in c2; do timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_first_${wl}.json; cat gpurun_out/bench_first_${wl}.json | cut -c1-600; done
