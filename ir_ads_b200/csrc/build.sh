#!/usr/bin/env bash
# Builds ir_ads_b200/libmsda_b200.so (the C-ABI library of include/msda.h) for sm_100a, in tree.
# nvcc cross-compiles without a GPU; the .so is git-ignored but ships to the GPU box with gpurun.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="${here}/../libmsda_b200.so"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xptxas -v -shared -Xcompiler -fPIC \
  -o "${out}" "${here}/msda_capi.cu" > "${here}/build.log" 2>&1 || { cat "${here}/build.log"; exit 1; }
echo "built ${out}"
