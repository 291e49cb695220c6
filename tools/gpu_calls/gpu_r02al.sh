#!/usr/bin/env bash
# Round-2 GPU call AL: fixed-point contributions from the exact fp64 product (one DFMA with a 1.5*2^52 addend instead of
# FMUL + F2I.S64, csrc/msda_coords.cuh det_contrib) -- whole GPU suite + smoke with the new library, then the deterministic
# backward against a variant build with the previous definition (build/variants/lib_detf32.so).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02al.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02al.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02al.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02al.log"
for v in detf32 product; do
  echo "== $v" | tee -a "$out/sweep_det_r02al.log"
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 python tools/sweep.py --iters 10 --det --workloads cfg2,cfg5,cfg3_f32 2>&1 | grep -v "^\[" | cut -c1-125 | tee -a "$out/sweep_det_r02al.log"
done
tail -3 "$out/pytest_r02al.log"; tail -2 "$out/smoke_r02al.log"
