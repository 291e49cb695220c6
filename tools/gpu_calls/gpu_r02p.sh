#!/usr/bin/env bash
# Round-2 GPU call P: backward write-out of grad_sampling_loc / grad_attn_weight in full lines (parked in the records)
# against per-point stores; 16-byte alignment precondition; whole parity suite.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02p.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02p.log"
{
for v in base stco base stco; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5,cfg3 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_stco_r02p.log" 2>&1
tail -4 "$out/pytest_r02p.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_stco_r02p.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:12s} {d['workload']:5s} {d['dist']:6s} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
