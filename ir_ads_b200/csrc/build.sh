#!/usr/bin/env bash
# Builds ir_ads_b200/libmsda_b200.so (the C-ABI library of include/msda.h) for sm_100a, in tree.
# nvcc cross-compiles without a GPU; the .so is git-ignored but ships to the GPU box with gpurun.
# Two translation units, compiled in parallel and only when one of their sources is newer than the object.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="${here}/../libmsda_b200.so"
obj="${here}/../../build/obj"
mkdir -p "${obj}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -Xcompiler -fPIC"
hdrs="${here}/msda_coords.cuh ${here}/msda_fast.cuh ${here}/../../include/msda.h"
compile() {   # compile <name> <extra deps...>
  local name="$1"; shift
  local src="${here}/${name}.cu" o="${obj}/${name}.o" stale=0
  for f in "${src}" ${hdrs} "$@" "${BASH_SOURCE[0]}"; do
    if [[ ! -f "${o}" || "${f}" -nt "${o}" ]]; then stale=1; fi
  done
  if [[ ${stale} == 1 ]]; then
    "${NVCC}" ${FLAGS} -c -o "${o}" "${src}" > "${here}/build_${name}.log" 2>&1 || { cat "${here}/build_${name}.log"; return 1; }
  fi
}
compile msda_capi "${here}/msda_generic.cuh" "${here}/msda_det.cuh" & p1=$!
compile msda_coarse_launch "${here}/msda_coarse.cuh" & p2=$!
wait ${p1}; wait ${p2}
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${obj}/msda_capi.o" "${obj}/msda_coarse_launch.o"
cat "${here}"/build_msda_*.log > "${here}/build.log" 2>/dev/null || true
echo "built ${out}"
