#!/usr/bin/env bash
# Round-2 GPU call I: 16-byte forward record (MSDA_FWD_REC16) against the 20-byte one: results + A/B timing
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 300 python tools/compare_variant.py build/variants/lib_rec16.so cfg2,cfg5 > "$out/compare_rec16_r02i.log" 2>&1; echo "compare exit $?" >> "$out/compare_rec16_r02i.log"
{
for rep in 1 2; do
for v in slim rec16; do
  echo "== $v (rep $rep)"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
done
} > "$out/sweep_rec16_r02i.log" 2>&1
tail -30 "$out/compare_rec16_r02i.log"; cat "$out/sweep_rec16_r02i.log"
