#!/usr/bin/env bash
# Round-2 GPU call A: first timings of the folding backward (product library and register variants), the parity
# suite, one ncu capture of the folding kernel.  Every stage under its own timeout, unbuffered logs.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu_r02a.csv" 2>&1
S="timeout 200 python -u tools/sweep.py --iters 15"
{
echo "== product library"
$S --workloads cfg2 --dists model,test --flags 0,8192,4096,32,4
$S --workloads cfg4,cfg5,cfg2_bf16 --flags 8192,4096
$S --workloads cfg3,cfg3_f32 --flags 0
echo "== fold variants (slim)"
for v in fold2 fold3; do
  echo "-- $v"
  MSDA_B200_LIB=build/variants/lib_$v.so $S --workloads cfg2 --dists model,test --flags 4096
  MSDA_B200_LIB=build/variants/lib_$v.so $S --workloads cfg5 --flags 4096
done
} > "$out/sweep_r02a.log" 2>&1
timeout 1100 python -u -m pytest tests -m gpu -x -q -s --timeout 300 --timeout-method=thread > "$out/pytest_r02a.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02a.log"
PROF="python bench.py --steps 1 --warmup 3 --layers 1 --no-cpu-baseline --no-e2e --no-ref-cuda --flags 4096"
timeout 120 $PROF > "$out/prof_plain_r02a.log" 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_fold -s 3 -c 1 -f -o "$out/prof_fold_r02a" $PROF > "$out/ncu_fold_r02a.log" 2>&1
tail -5 "$out/pytest_r02a.log"; cat "$out/sweep_r02a.log"
