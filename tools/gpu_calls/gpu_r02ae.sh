#!/usr/bin/env bash
# Round-2 GPU call AE: deterministic sorted path -- multiply-high divisions in the count pass (fd), cell reduce at 4 CTAs
# per SM (fdm4), and the fixed-point scale's two maxima taken per warp by the entry-filing backward (product) --
# parity first (bit equality with the fixed-point red path), then A/B timing against the previous library (head).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "determin or det_ or cfg5" > "$out/pytest_det_r02ae.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_det_r02ae.log"
tail -3 "$out/pytest_det_r02ae.log"
for v in head fd fdm4 product head product; do
  echo "== $v" | tee -a "$out/sweep_det_r02ae.log"
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 python tools/sweep.py --iters 15 --det --workloads cfg2,cfg5 2>&1 | grep -v "^\[" | tee -a "$out/sweep_det_r02ae.log"
done
