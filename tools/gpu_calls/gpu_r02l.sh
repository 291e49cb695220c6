#!/usr/bin/env bash
# Round-2 GPU call L (2 GPUs): the driver's multi-GPU launch of bench.py after the round's last edits -- both arms -- and
# the training-step mode, strong scaling, at N=2.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > "$out/bench_ref_n2_r02l.json" 2> "$out/bench_ref_n2_r02l.err"; echo "ref exit $?"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > "$out/bench_n2_r02l.json" 2> "$out/bench_n2_r02l.err"; echo "bench exit $?"
timeout 300 $TR bench.py --gpus 2 --workload cfg4 --mode train --scaling strong --total-batch 16 --steps 10 --warmup 3 > "$out/bench_train_n2_r02l.json" 2> "$out/bench_train_n2_r02l.err"; echo "train exit $?"
timeout 300 $TR bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda > "$out/bench_cfg5_n2_r02l.json" 2> "$out/bench_cfg5_n2_r02l.err"; echo "cfg5 exit $?"
for f in bench_ref_n2 bench_n2 bench_train_n2 bench_cfg5_n2; do echo "== $f"; head -c 700 "$out/${f}_r02l.json"; echo; tail -2 "$out/${f}_r02l.err"; done
