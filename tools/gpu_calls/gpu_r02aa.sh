#!/usr/bin/env bash
# Round-2 GPU call AA: the training-step mode with the final kernels at N=1 -- fp32, fp32 with TF32 projections (the default of
# the torch 1.10 the reference pins), bf16 autocast.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
: > "$out/bench_train_r02aa.jsonl"
for cfg in "" "--tf32" "--amp"; do
  timeout 300 python bench.py --workload cfg4 --mode train --scaling strong --total-batch 16 --steps 10 --warmup 3 $cfg >> "$out/bench_train_r02aa.jsonl" 2>> "$out/bench_train_r02aa.err"
done
python - <<'PY'
import json
for line in open("gpurun_out/bench_train_r02aa.jsonl"):
    if not line.startswith("{"): continue
    d=json.loads(line); print(d["dtype"], round(d["ms_per_step"],3), "ms/step", d["config"]["execution"])
PY
