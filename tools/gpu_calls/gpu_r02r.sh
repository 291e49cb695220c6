#!/usr/bin/env bash
# Round-2 GPU call R (8 GPUs): the final kernels at N=8 -- cfg2 weak (the metric), cfg5 deterministic weak, cfg4 training step strong.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda > "$out/bench_n8_r02r.json" 2> "$out/bench_n8_r02r.err"; echo "bench exit $?"
timeout 300 $TR bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda > "$out/bench_cfg5_n8_r02r.json" 2> "$out/bench_cfg5_n8_r02r.err"; echo "cfg5 exit $?"
timeout 300 $TR bench.py --gpus 8 --workload cfg4 --mode train --scaling strong --total-batch 16 --steps 10 --warmup 3 > "$out/bench_train_n8_r02r.json" 2> "$out/bench_train_n8_r02r.err"; echo "train exit $?"
for f in bench_n8 bench_cfg5_n8 bench_train_n8; do echo "== $f"; grep "^{" "$out/${f}_r02r.json" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3),'G pts/s', round(d['ms_per_step'],3),'ms/step', d.get('timed_regions'), (d.get('e2e') or {}).get('value'))"; tail -1 "$out/${f}_r02r.err"; done
