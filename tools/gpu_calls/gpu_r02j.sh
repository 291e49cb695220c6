#!/usr/bin/env bash
# Round-2 GPU call J: unclamped records + validity-bit predicated corner loads + 16-byte forward record.
# Whole parity suite on the product build, experiment tests on the experiment build, A/B timing of the variants.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q -s --timeout 300 --timeout-method=thread > "$out/pytest_r02j.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02j.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological" > "$out/pytest_exp_r02j.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02j.log"
{
for v in slim predg1 pred20 predg2m4 predg2m5 slim predg1; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5,cfg3 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_pred_r02j.log" 2>&1
tail -4 "$out/pytest_r02j.log"; tail -3 "$out/pytest_exp_r02j.log"; cat "$out/sweep_pred_r02j.log"
