#!/usr/bin/env bash
# Round-2 GPU call Z: pre-reduction of the scatter inside a row (weights of corners that land on a pixel an earlier point of
# the same row and level already scatters to are handed over; no red for them) -- the whole suite against a FULL build with
# -DMSDA_BWD_MERGE=1, then A/B timing of slim builds.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
MSDA_B200_LIB=build/variants/lib_mergefull.so timeout 1200 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02z.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02z.log"
{
for v in nomerge merge nomerge merge; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5,cfg3,cfg3_f32,cfg4 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_merge_r02z.log" 2>&1
tail -4 "$out/pytest_r02z.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_merge_r02z.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:10s} {d['workload']:9s} {d['dist']:6s} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
