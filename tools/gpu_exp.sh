#!/usr/bin/env bash
# experiment: persistent TILED order with 256-thread CTAs (6 per SM) vs the 1024-thread one vs LINEAR
set -u
out=gpurun_out
mkdir -p "$out"
S="timeout 300 python tools/sweep.py --iters 10"
$S --workloads cfg2 --flags 0,8 2>&1 | tail -2
MSDA_EXP_TILED_256=1 $S --workloads cfg2,cfg5 --flags 8 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "pytorch_named or bf16 or tiled" 2>&1 | tail -3
