#!/usr/bin/env bash
# Builds a slim kernel-variant library into build/variants/lib_<name>.so (never the product .so):
#   bash tools/build_variant.sh <name> [-DMSDA_FWD_THREADS=128 ...]
# -DMSDA_EXP_SLIM instantiates only D = 32, P in {4, 8}, LINEAR / STRIP orders (about 20 s per build).
# Time it on the GPU box with tools/variant_sweep.sh "<name> ..." (MSDA_B200_LIB selects the library).
set -euo pipefail
name="$1"; shift
cd "$(dirname "$0")/../ir_ads_b200/csrc"
mkdir -p ../../build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DMSDA_EXP_SLIM "$@" -Xptxas -v -shared -Xcompiler -fPIC \
  -o "../../build/variants/lib_${name}.so" msda_capi.cu msda_coarse_launch.cu > "/tmp/variant_${name}.log" 2>&1 || { grep -i error "/tmp/variant_${name}.log"; exit 1; }
echo "built build/variants/lib_${name}.so (ptxas log: /tmp/variant_${name}.log)"
