#!/usr/bin/env bash
# ncu --set full of the kernels behind the other named configs: cfg3 (decoder, bf16 value) and cfg5 (5 levels,
# 8 points, deterministic backward).  Each capture only after the same command exited 0 without ncu.
set -u
out=gpurun_out
mkdir -p "$out"
P3="python tools/sweep.py --workloads cfg3 --iters 1"
P5="python tools/sweep.py --workloads cfg5 --iters 1 --det"
$P3 > "$out/ncu_other_plain3.log" 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_|det_" -s 9 -c 3 -f -o "$out/prof_cfg3" $P3 > "$out/ncu_cfg3.log" 2>&1
$P5 > "$out/ncu_other_plain5.log" 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_|det_" -s 39 -c 13 -f -o "$out/prof_cfg5det" $P5 > "$out/ncu_cfg5det.log" 2>&1
# the reports are too large to travel back (64 MiB limit on gpurun_out/): summarise here, keep the text only
: > "$out/ncu_other_summary.txt"
for r in prof_cfg3 prof_cfg5det; do
  echo "######## $r" >> "$out/ncu_other_summary.txt"
  python tools/ncu_summary.py "$out/$r.ncu-rep" "$out/ncu_other_summary.txt" > /dev/null 2>&1
  rm -f "$out/$r.ncu-rep"
done
wc -l "$out/ncu_other_summary.txt"
