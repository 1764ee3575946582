#!/bin/bash
# diagnostics: microbench, pipeline trace, ncu full captures of the lattice kernel per workload
TAG=${1:-d1}
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
./profiles/microbench/ubench > gpurun_out/ubench_${TAG}.txt 2>&1
timeout 120 python scratch/trace.py c2 > gpurun_out/trace_c2_${TAG}.txt 2>&1
for wl in c2 c3 c5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lattice -s 3 -c 1 -o gpurun_out/prof_lattice_${wl}_${TAG} -f python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${wl}_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${wl}_${TAG}.log
done
ls -la gpurun_out | tail
