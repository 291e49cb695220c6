#!/usr/bin/env bash
# Round-2 GPU call B: guarded (every stage under its own timeout, unbuffered logs) after call A hung.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
S="python -u tools/sweep.py --iters 10"
{
echo "== product library, no fold"; timeout 150 $S --workloads cfg2 --dists model --flags 8192; echo "rc $?"
echo "== product library, fold";    timeout 150 $S --workloads cfg2 --dists model,test --flags 4096; echo "rc $?"
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
echo "== fold3 variant";            MSDA_B200_LIB=build/variants/lib_fold3.so timeout 150 $S --workloads cfg2 --dists model,test --flags 4096; echo "rc $?"
} > "$out/sweep_r02b.log" 2>&1
timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -v --timeout 120 --timeout-method=thread -k "fold or row_orders or tile2d or bookkeeping" > "$out/pytest_fold_r02b.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_fold_r02b.log"
timeout 1200 python -u -m pytest tests -m gpu -x -v -s --timeout 240 --timeout-method=thread > "$out/pytest_r02b.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02b.log"
tail -15 "$out/pytest_fold_r02b.log"; tail -15 "$out/pytest_r02b.log"; cat "$out/sweep_r02b.log"
