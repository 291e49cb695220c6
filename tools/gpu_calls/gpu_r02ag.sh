#!/usr/bin/env bash
# Round-2 GPU call AG: deterministic sorted path after call AF's finding (the cell reduce is sensitive to its launch
# bound: keep the plain one) -- candidate product = multiply-high divisions in the count pass + per-warp maxima from the
# entry-filing backward; variants: without those maxima (noamax), cell reduce bounded to 2 / 3 CTAs per SM (m2, m3).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "determin or det_ or cfg5" > "$out/pytest_det_r02ag.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_det_r02ag.log"
tail -3 "$out/pytest_det_r02ag.log"
for v in head noamax product m2 m3 head product; do
  echo "== $v" | tee -a "$out/sweep_det_r02ag.log"
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 python tools/sweep.py --iters 15 --det --workloads cfg2,cfg5,cfg3_f32 2>&1 | grep -v "^\[" | cut -c1-120 | tee -a "$out/sweep_det_r02ag.log"
done
