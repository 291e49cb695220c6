#!/usr/bin/env bash
# Round-2 GPU call T: row orders re-measured on the final kernels (LINEAR = 4, STRIP = 16, TILE2D = 32 for both passes);
# module-, layer- and stack-level timings with the final library.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 300 python tools/sweep.py --workloads cfg2,cfg5 --dists model,test --flags 0,4,16,32 --iters 20 > "$out/sweep_orders_r02t.log" 2>&1
timeout 300 python tools/bench_module.py > "$out/bench_module_r02t.json" 2> "$out/bench_module_r02t.err"
timeout 300 python tools/bench_layer.py > "$out/bench_layer_r02t.json" 2> "$out/bench_layer_r02t.err"
python - <<'PY'
import json
for line in open("gpurun_out/sweep_orders_r02t.log"):
    try: d=json.loads(line)
    except Exception: continue
    print(f"{d['workload']:5s} {d['dist']:6s} flags {d['flags']:3d} fwd {d['fwd_ms']:.4f} bwd {d['bwd_ms']:.4f}")
PY
tail -c 2500 "$out/bench_module_r02t.json"; echo; tail -c 1500 "$out/bench_layer_r02t.json"
