#!/usr/bin/env bash
# Multi-GPU measurements of round 2 at N GPUs (run under `gpurun --gpus N`): weak and strong scaling of the core
# operator, the batch-sharded training step (BASELINE config 4), the deterministic stress config (BASELINE config 5).
# Usage: bash tools/scaling_r02.sh N     -> gpurun_out/scale_r02_n<N>.jsonl (one bench.py JSON line per configuration)
set -u
N="${1:-1}"
out=gpurun_out
mkdir -p "$out"
f="$out/scale_r02_n${N}.jsonl"
: > "$f"
export PYTHONUNBUFFERED=1
run() {
  if [[ "$N" == 1 ]]; then timeout 300 python bench.py --gpus 1 "$@" >> "$f" 2>> "$out/scale_r02_n${N}.err"
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus "$N" "$@" >> "$f" 2>> "$out/scale_r02_n${N}.err"
  fi
  echo "rc $? : $*" >> "$out/scale_r02_n${N}.err"
}
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda"
run $Q --no-e2e                                                        # cfg2 weak (8 images per GPU): the headline workload
run $Q --no-e2e --scaling strong --total-batch 8                       # cfg2 strong, 8 images in the job
run $Q --no-e2e --scaling strong --total-batch 16                      # cfg2 strong, 16 images in the job
run $Q --no-e2e --workload cfg5                                        # cfg5 weak: 8 images per GPU, deterministic backward (B=64 at N=8)
run $Q --workload cfg4 --mode train --scaling strong --total-batch 16  # cfg4 training step, fp32
run $Q --workload cfg4 --mode train --scaling strong --total-batch 16 --no-graph
run $Q --workload cfg4 --mode train --scaling strong --total-batch 16 --amp
python - "$f" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try: d = json.loads(line)
    except Exception: continue
    c = d["config"]
    print(f"N={d['n_gpus']} {d['scaling']:6s} {c['workload'][:34]:34s} B/gpu={c['batch_per_gpu']:2d} {d['dtype'][:4]} {d['ms_per_step']:9.3f} ms/step {d['value']/1e9:8.3f} Gpts/s  regions {d['timed_regions']['ms_per_step']} {c.get('execution','')[:40]}")
PY
tail -12 "$out/scale_r02_n${N}.err"
