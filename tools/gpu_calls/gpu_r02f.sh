#!/usr/bin/env bash
# Round-2 GPU call F: TILE2D with image-major tile order at 128 / 256-thread CTAs; the folding backward with the same order
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
{
for v in t128 t256; do
  echo "== $v"
  MSDA_B200_LIB=build/variants/lib_$v.so timeout 200 python -u tools/sweep.py --iters 15 --workloads cfg2 --dists model,test --flags 0,32
done
echo "== expslim (fold, 3 CTAs/SM)"
MSDA_B200_LIB=build/variants/lib_expslim.so timeout 200 python -u tools/sweep.py --iters 15 --workloads cfg2 --dists model,test --flags 4096
} > "$out/sweep_r02f2.log" 2>&1
cat "$out/sweep_r02f2.log"
