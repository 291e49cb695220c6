#!/usr/bin/env bash
# experiment: forward row orders (0 LINEAR, 16 STRIP, 1040 STRIP head-major, 32 TILE2D, 8 TILED) + ncu of two of them
set -u
out=gpurun_out
mkdir -p "$out"
S="timeout 300 python tools/sweep.py --iters 10"
$S --workloads cfg2 --dists model,test --flags 0,16,1040,32,8 2>&1 | tee "$out/exp_fwd_orders_cfg2.jsonl"
$S --workloads cfg3,cfg5,cfg2_bf16 --flags 0,16,1040 2>&1 | tee "$out/exp_fwd_orders_other.jsonl"
for fl in 0 1040; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -f -o "$out/prof_fwd_flags$fl" \
    python tools/sweep.py --workloads cfg2 --flags $fl --iters 2 > "$out/ncu_fwd_flags$fl.log" 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q > "$out/pytest_full.log" 2>&1; echo "exit $?" >> "$out/pytest_full.log"
tail -5 "$out/pytest_full.log"
