#!/usr/bin/env bash
# Builds experiment variants of the library into build/variants/ (never the product .so):
#   no_red     backward without the grad_value scatter   -> cost of the gather + reductions alone
#   no_gather  backward without the value gather          -> cost of the scatter alone
set -euo pipefail
cd "$(dirname "$0")/../ir_ads_b200/csrc"
mkdir -p ../../build/variants
for v in NO_RED NO_GATHER; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DMSDA_EXP_$v -shared -Xcompiler -fPIC \
    -o ../../build/variants/lib_$(echo $v | tr A-Z a-z).so msda_capi.cu > /tmp/ablate_$v.log 2>&1 &
done
wait
ls -la ../../build/variants/
