#!/usr/bin/env bash
# Round-2 GPU call AB: DCNv3 backward with the row-local pre-reduction of the scatter weights (taps of a convolution grid
# overlap heavily) -- DCNv3 tests, then A/B against a full build without it.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_dcnv3.py -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_dcn_r02ab.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_dcn_r02ab.log"
for rep in 1 2; do
  echo "== no merge (rep $rep)"; MSDA_B200_LIB=build/variants/lib_dcnnomerge.so timeout 300 python tools/bench_dcnv3.py 2>&1 | grep "^{"
  echo "== merge (rep $rep)"; timeout 300 python tools/bench_dcnv3.py 2>&1 | grep "^{"
done > "$out/bench_dcn_r02ab.log" 2>&1
tail -3 "$out/pytest_dcn_r02ab.log"; python - <<'PY'
import json
cur=None
for line in open("gpurun_out/bench_dcn_r02ab.log"):
    if line.startswith("=="): cur=line.strip(); continue
    try: d=json.loads(line)
    except Exception: print(line.strip()); continue
    print(f"{cur:22s} {str(d['shape']):24s} fwd {d['b200_fwd_ms']:.4f} bwd {d['b200_bwd_ms']:.4f}  ref fwd {d.get('ref_fwd_ms')} bwd {d.get('ref_bwd_ms')}")
PY
