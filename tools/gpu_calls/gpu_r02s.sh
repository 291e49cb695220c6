#!/usr/bin/env bash
# Round-2 GPU call S: degenerate-pyramid parity test; backward occupancy (launch bounds for 8 / 10 / 12 CTAs per SM).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "degenerate or misaligned or non_finite" > "$out/pytest_r02s.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02s.log"
{
for v in base bw5 bw6 base bw5; do
  echo "== $v"
  MSDA_B200_LIB="build/variants/lib_${v}.so" timeout 300 python tools/sweep.py --workloads cfg2,cfg5,cfg3 --dists model,test --iters 30 2>&1 | grep -v "^\["
done
} > "$out/sweep_bwocc_r02s.log" 2>&1
tail -3 "$out/pytest_r02s.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/sweep_bwocc_r02s.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:12s} {d['workload']:5s} {d['dist']:6s} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
