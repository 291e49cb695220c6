#!/usr/bin/env bash
# Round-2 GPU call O: persistent forward CTAs (resident CTAs stride over the row groups; optional prefetch of the next
# row's locations before the gather) against one CTA per row group.  Results compared, then A/B timing.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
for v in pers1 pers2; do timeout 200 python tools/compare_variant.py build/variants/lib_$v.so cfg2 2>&1 | tail -2; done > "$out/compare_pers_r02o.log" 2>&1
{
for v in base pers1 pers2 pers1x2 pers1m5 pers2m5 base; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5,cfg3 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_pers_r02o.log" 2>&1
cat "$out/compare_pers_r02o.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_pers_r02o.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:12s} {d['workload']:5s} {d['dist']:6s} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
