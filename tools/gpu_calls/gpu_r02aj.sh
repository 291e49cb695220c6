#!/usr/bin/env bash
# Round-2 GPU call AJ (2 GPUs): the driver's launch line at N=2 with the final library -- default bench, the reference
# arm (rank 0 alone works), and cfg5 (deterministic, its workspace layout changed in call AG).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > "$out/bench_n2_r02aj.json" 2> "$out/bench_n2_r02aj.err"; echo "bench exit $?" >> "$out/bench_n2_r02aj.err"
timeout 300 $TR --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > "$out/bench_ref_n2_r02aj.json" 2> "$out/bench_ref_n2_r02aj.err"; echo "ref exit $?" >> "$out/bench_ref_n2_r02aj.err"
timeout 300 $TR --master-port 29613 bench.py --gpus 2 --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > "$out/bench_cfg5_n2_r02aj.json" 2> "$out/bench_cfg5_n2_r02aj.err"; echo "cfg5 exit $?" >> "$out/bench_cfg5_n2_r02aj.err"
for f in bench_n2 bench_ref_n2 bench_cfg5_n2; do python -c "
import json,sys
ls=[l for l in open('$out/${f}_r02aj.json').read().splitlines() if l.startswith('{')]
d=json.loads(ls[-1]); print('$f', len(ls), 'line(s)', d.get('impl','b200'), d['n_gpus'], d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))"; tail -1 "$out/${f}_r02aj.err"; done
