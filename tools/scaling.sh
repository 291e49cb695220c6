#!/usr/bin/env bash
# 1/2/4/8-GPU scaling of bench.py on one box (run under `gpurun --gpus 8`). Writes gpurun_out/scale_<tag>.jsonl
set -u
tag="${1:-r01}"
out="gpurun_out/scale_${tag}.jsonl"
: > "$out"
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline >> "$out" 2>> "gpurun_out/scale_${tag}.err"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus "$n" --steps 10 --warmup 3 --no-cpu-baseline >> "$out" 2>> "gpurun_out/scale_${tag}.err"
  fi
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
base = rows[0]["value"]
for r in rows:
    print(f'N={r["n_gpus"]}  value={r["value"]/1e9:8.3f} Gpts/s  x{r["value"]/base:5.2f}  ms/step={r["ms_per_step"]:.3f}  e2e={r["e2e"]["value"]/1e9:.3f} Gpts/s')
PY
