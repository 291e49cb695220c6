#!/usr/bin/env bash
# Round-2 GPU call E: deterministic path with entries emitted by the backward kernel (tests + A/B timing), generic DCNv3
# composition tests, single-GPU batch sweep of the training-step mode (what limits strong scaling), its launch list.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py tests/test_dcnv3.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "determin or dcnv3 or random_problem or cfg5_det or kernels_actually" > "$out/pytest_r02e.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02e.log"
{
echo "== deterministic backward: entries from the backward kernel (flags 0) vs separate fill pass (flags 16384)"
timeout 300 python -u tools/sweep.py --iters 10 --det --workloads cfg2,cfg5,cfg4,cfg3_f32 --flags 0,16384
} > "$out/sweep_det_r02e.log" 2>&1
for b in 2 4 8 16; do
  timeout 200 python bench.py --workload cfg4 --mode train --scaling strong --total-batch $b --steps 10 --warmup 3 >> "$out/train_batch_sweep_r02e.jsonl" 2>> "$out/train_batch_sweep_r02e.err"
done
PROF="python bench.py --workload cfg4 --mode train --scaling strong --total-batch 2 --steps 1 --warmup 3 --regions 1 --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$out/launches_train_b2_r02e.csv" $PROF > "$out/ncu_train_b2_r02e.log" 2>&1
PROF="python bench.py --workload cfg4 --mode train --scaling strong --total-batch 16 --steps 1 --warmup 3 --regions 1 --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$out/launches_train_b16_r02e.csv" $PROF > "$out/ncu_train_b16_r02e.log" 2>&1
tail -4 "$out/pytest_r02e.log"; cat "$out/sweep_det_r02e.log"; python - <<'PY'
import json
for line in open("gpurun_out/train_batch_sweep_r02e.jsonl"):
    if line.startswith("{"):
        d = json.loads(line); print(d["config"]["batch_per_gpu"], d["ms_per_step"], d["ms_per_step"] / d["config"]["batch_per_gpu"])
PY
