#!/usr/bin/env bash
# Round-2 GPU call X: bf16 forward with 4 instead of 8 channels per lane on the small decoder problem (cfg3) and on cfg2_bf16.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
{
for v in base cpl4 base cpl4; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg3,cfg2_bf16 --dists model,test --iters 40 2>&1 | grep -v "^\["
done
} > "$out/sweep_cpl4_r02x.log" 2>&1
python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_cpl4_r02x.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:10s} {d['workload']:9s} {d['dist']:6s} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
