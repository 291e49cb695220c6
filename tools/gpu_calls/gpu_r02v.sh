#!/usr/bin/env bash
# Round-2 GPU call V (4 GPUs): the final kernels at N=4 (cfg2 weak, the metric) -- completes the 1/2/4/8 line of the final library.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda > "$out/bench_n4_r02v.json" 2> "$out/bench_n4_r02v.err"; echo "bench exit $?"
grep "^{" "$out/bench_n4_r02v.json" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3),'G pts/s', round(d['ms_per_step'],3),'ms/step', d.get('timed_regions'), (d.get('e2e') or {}).get('value'))"
