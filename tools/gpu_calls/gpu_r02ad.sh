#!/usr/bin/env bash
# Round-2 GPU call AD: deterministic sorted path with (1) the two amax passes folded into the backward kernel that files
# the entries, (2) multiply-high divisions in the count pass, (3) cell reduce at 3 / 4 CTAs per SM -- parity first
# (bit equality with the fixed-point red path), then A/B timing against the previous library (build/variants/lib_head.so).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "determin or det_ or cfg5" > "$out/pytest_det_r02ad.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_det_r02ad.log"
tail -3 "$out/pytest_det_r02ad.log"
for v in head product rminb4; do
  echo "== $v" | tee -a "$out/sweep_det_r02ad.log"
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 python tools/sweep.py --iters 15 --det --workloads cfg2,cfg5 2>&1 | grep -v "^\[" | tee -a "$out/sweep_det_r02ad.log"
done
