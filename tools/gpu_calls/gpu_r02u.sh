#!/usr/bin/env bash
# Round-2 GPU call U: LINEAR order with a dispatch swizzle (row groups 148 apart in launch order are neighbours in memory,
# chunks of 148 x K groups), forward and LINEAR backward; results compared with the product library first.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 200 python tools/compare_variant.py build/variants/lib_sw150.so cfg2 2>&1 | tail -2 > "$out/compare_sw_r02u.log"
{
for v in base sw24 sw75 sw150 base; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5 --dists model,test --flags 0,4 --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_sw_r02u.log" 2>&1
cat "$out/compare_sw_r02u.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_sw_r02u.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:10s} {d['workload']:5s} {d['dist']:6s} flags {d['flags']} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
