#!/usr/bin/env bash
# Builds experiment variants of the library into build/variants/ (never the product .so):
#   no_red     backward without the grad_value scatter   -> cost of the gather + reductions alone
#   no_gather  backward without the value gather          -> cost of the scatter alone
#   knobs      the product kernels plus the MSDA_EXP_* environment knobs of msda_capi.cu (occupancy / carve-out /
#              coarse-kernel-skipped studies of profiles/r01q_experiments.txt); use with
#              MSDA_B200_LIB=build/variants/lib_knobs.so python tools/sweep.py ...
set -euo pipefail
cd "$(dirname "$0")/../ir_ads_b200/csrc"
mkdir -p ../../build/variants
for v in NO_RED NO_GATHER; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DMSDA_EXP_$v -shared -Xcompiler -fPIC \
    -o ../../build/variants/lib_$(echo $v | tr A-Z a-z).so msda_capi.cu msda_coarse_launch.cu > /tmp/ablate_$v.log 2>&1 &
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DMSDA_EXPERIMENTS -shared -Xcompiler -fPIC \
  -o ../../build/variants/lib_knobs.so msda_capi.cu msda_coarse_launch.cu > /tmp/ablate_knobs.log 2>&1 &
wait
ls -la ../../build/variants/
