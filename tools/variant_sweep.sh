#!/usr/bin/env bash
# Times kernel-variant builds (build/variants/lib_*.so, see MSDA_EXP_SLIM in msda_capi.cu) with tools/sweep.py.
# Usage (on the GPU box): bash tools/variant_sweep.sh "slim early t128" [workloads] [flags]
set -u
variants="${1:-slim}"
workloads="${2:-cfg2}"
flags="${3:-0}"
mkdir -p gpurun_out
for v in $variants; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" python tools/sweep.py --workloads "$workloads" --flags "$flags" --iters 20 2>&1 | grep -v "^\[" | tee -a "gpurun_out/variants.log"
done
