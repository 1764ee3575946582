#!/bin/bash
# Round-2 evidence, run on the GPU box: ncu launch list of the bench command, one `ncu --set full` capture per kernel
# (raw-metric CSV + per-source-line CSV exported on the box; the .ncu-rep of the headline kernel is kept), and
# compute-sanitizer logs.  Everything lands in gpurun_out/r02/.
#   bash tools/capture_profiles.sh [ncu|san|all]
set -u
what=${1:-all}
out=gpurun_out/r02
mkdir -p $out
NCU="ncu --clock-control none"
if [ "$what" = "ncu" ] || [ "$what" = "all" ]; then
  $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
  cap() {  # name, kernel regex, launch-skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU --set full --import-source on -k "regex:$regex" -s $skip -c 1 -f -o $out/$name "$@" > $out/$name.log 2>&1
    if [ -f $out/$name.ncu-rep ]; then
      ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
      ncu -i $out/$name.ncu-rep --page source --csv --print-source cuda,sass > $out/$name.source.csv 2>/dev/null
      python tools/ncu_summary.py $out/$name.raw.csv > $out/$name.summary.txt 2>&1
      python tools/ncu_lines.py $out/$name.source.csv 30 > $out/$name.lines.txt 2>&1
      rm -f $out/$name.source.csv
      [ "$name" = "ncu_wave_c2" ] || rm -f $out/$name.ncu-rep
    fi
  }
  cap ncu_wave_c2      ctc_wave_kernel      2 python tools/run_one.py c2 4
  cap ncu_wave_c1      ctc_wave_kernel      2 python tools/run_one.py c1 4
  cap ncu_sweep_c3     ctc_sweep_kernel     2 python tools/run_one.py c3 4
  cap ncu_sweep_c5     ctc_sweep_kernel     1 python tools/run_one.py c5 3 512
  cap ncu_rowstats_c4  ctc_row_stats        2 python tools/run_one.py c4 4
  cap ncu_general_c4   ctc_fused_kernel     2 python tools/run_one.py c4 4
  cap ncu_grad_c4      ctc_grad_kernel      2 python tools/run_one.py c4 4
  cap ncu_reduce_c2    ctc_loss_reduce      2 python tools/run_one.py c2 4
  cap ncu_argmax_c4    ctc_argmax_kernel    1 python tools/run_one.py c4 2 0 --greedy
  cap ncu_collapse_c4  ctc_collapse_kernel  1 python tools/run_one.py c4 2 0 --greedy
  cap ncu_scale_c2     ctc_scale_rows       1 python tools/run_one.py c2 2 0 --scale
  cap ncu_viterbi_c2   ctc_viterbi_kernel   1 python tools/run_one.py c2 2 0 --align
  cap ncu_noblank_c2   ctc_noblank_kernel   1 python tools/run_one.py c2 2 0 --noblank
  cap ncu_general_c1   ctc_fused_kernel     2 python tools/run_one.py c1 4
  cap ncu_beam_c2      ctc_beam_kernel      1 python tools/beam_probe.py c2 100 1
fi
if [ "$what" = "beam" ]; then         # the prefix-beam-search kernel (added late in round 2)
  cap() {
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU --set full --import-source on -k "regex:$regex" -s $skip -c 1 -f -o $out/$name "$@" > $out/$name.log 2>&1
    if [ -f $out/$name.ncu-rep ]; then
      ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
      ncu -i $out/$name.ncu-rep --page source --csv --print-source cuda,sass > $out/$name.source.csv 2>/dev/null
      python tools/ncu_summary.py $out/$name.raw.csv > $out/$name.summary.txt 2>&1
      python tools/ncu_lines.py $out/$name.source.csv 30 > $out/$name.lines.txt 2>&1
      rm -f $out/$name.source.csv $out/$name.ncu-rep $out/$name.raw.csv
    fi
  }
  cap ncu_beam_c2      ctc_beam_kernel      1 python tools/beam_probe.py c2 100 1
  cap ncu_beam_c4      ctc_beam_kernel      1 python tools/beam_probe.py c4 100 1
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/beam_probe.py c1 100 1 > $out/san_${tool}_beam_c1.log 2>&1
    tail -2 $out/san_${tool}_beam_c1.log
  done
fi
if [ "$what" = "late" ]; then         # kernels changed late in round 2: K1 (symbol prefetch), K3 (posterior prefetch), sweep c3
  cap() {
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU --set full --import-source on -k "regex:$regex" -s $skip -c 1 -f -o $out/$name "$@" > $out/$name.log 2>&1
    if [ -f $out/$name.ncu-rep ]; then
      ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
      ncu -i $out/$name.ncu-rep --page source --csv --print-source cuda,sass > $out/$name.source.csv 2>/dev/null
      python tools/ncu_summary.py $out/$name.raw.csv > $out/$name.summary.txt 2>&1
      python tools/ncu_lines.py $out/$name.source.csv 30 > $out/$name.lines.txt 2>&1
      rm -f $out/$name.source.csv $out/$name.ncu-rep $out/$name.raw.csv
    fi
  }
  cap ncu_rowstats_c4  ctc_row_stats        2 python tools/run_one.py c4 4
  cap ncu_general_c4   ctc_fused_kernel     2 python tools/run_one.py c4 4
  cap ncu_grad_c4      ctc_grad_kernel      2 python tools/run_one.py c4 4
  cap ncu_sweep_c3     ctc_sweep_kernel     2 python tools/run_one.py c3 4
fi
if [ "$what" = "refresh" ]; then      # kernels changed after the first capture of the round
  NCUO=$NCU
  cap() {
    local name=$1 regex=$2 skip=$3; shift 3
    $NCUO --set full --import-source on -k "regex:$regex" -s $skip -c 1 -f -o $out/$name "$@" > $out/$name.log 2>&1
    if [ -f $out/$name.ncu-rep ]; then
      ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
      ncu -i $out/$name.ncu-rep --page source --csv --print-source cuda,sass > $out/$name.source.csv 2>/dev/null
      python tools/ncu_summary.py $out/$name.raw.csv > $out/$name.summary.txt 2>&1
      python tools/ncu_lines.py $out/$name.source.csv 30 > $out/$name.lines.txt 2>&1
      rm -f $out/$name.source.csv $out/$name.ncu-rep $out/$name.raw.csv
    fi
  }
  cap ncu_sweep_c3     ctc_sweep_kernel     2 python tools/run_one.py c3 4
  cap ncu_viterbi_c2   ctc_viterbi_kernel   1 python tools/run_one.py c2 2 0 --align
  cap ncu_noblank_c2   ctc_noblank_kernel   1 python tools/run_one.py c2 2 0 --noblank
  cap ncu_general_c1   ctc_fused_kernel     2 python tools/run_one.py c1 4
  cap ncu_general_c4   ctc_fused_kernel     2 python tools/run_one.py c4 4
fi
if [ "$what" = "san" ] || [ "$what" = "all" ]; then
  san() {  # tool, name, command...
    local tool=$1 name=$2; shift 2
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 "$@" > $out/san_${tool}_$name.log 2>&1
    tail -3 $out/san_${tool}_$name.log
  }
  san memcheck wave_c2 python tools/run_one.py c2 1 2
  san memcheck wave_c1 python tools/run_one.py c1 1
  san memcheck sweep_c3 python tools/run_one.py c3 1 160
  san memcheck general_c4_greedy_align python tools/run_one.py c4 1 3 --greedy --align
  for tool in racecheck synccheck; do
    san $tool wave_c2 python tools/run_one.py c2 1 2
    san $tool general_c4_greedy_align python tools/run_one.py c4 1 3 --greedy --align
  done
fi
ls -la $out
