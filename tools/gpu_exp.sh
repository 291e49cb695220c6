#!/usr/bin/env bash
# ncu --set full of the opt-in coarse-level pair (serial mode: flags 256|512 = 768): the coarse kernel and the main
# backward kernel with the coarse reds skipped
set -u
out=gpurun_out
mkdir -p "$out"
P="python tools/sweep.py --workloads cfg2 --flags 768 --iters 2"
$P > "$out/prof_coarse_plain.log" 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 6 -c 2 -f -o "$out/prof_coarse_pair" $P > "$out/ncu_coarse_pair.log" 2>&1
tail -3 "$out/ncu_coarse_pair.log"
