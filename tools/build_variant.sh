#!/usr/bin/env bash
# Builds a slim kernel-variant library into build/variants/lib_<name>.so (never the product .so):
#   bash tools/build_variant.sh <name> [-DMSDA_FWD_THREADS=128 ...]
#   FULL=1 bash tools/build_variant.sh exp -DMSDA_EXPERIMENTS     # every shape, plus the measured-slower experiments
# Without FULL=1, -DMSDA_EXP_SLIM instantiates only D = 32, P in {4, 8} of the plain operator (well under a minute).
# Time it on the GPU box with tools/variant_sweep.sh "<name> ..." (MSDA_B200_LIB selects the library).
set -euo pipefail
name="$1"; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "${root}/build/variants"
slim="-DMSDA_EXP_SLIM"; [[ "${FULL:-0}" == 1 ]] && slim=""
EXTRA_NVCC_FLAGS="${slim} $*" OUT="${root}/build/variants/lib_${name}.so" OBJ_DIR="${root}/build/variants/obj_${name}" \
  bash "${root}/ir_ads_b200/csrc/build.sh" > "/tmp/variant_${name}.log" 2>&1 || { grep -i "error" "/tmp/variant_${name}.log"; exit 1; }
echo "built build/variants/lib_${name}.so (ptxas log: build/variants/obj_${name}/build.log)"
