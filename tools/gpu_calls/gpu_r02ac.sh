#!/usr/bin/env bash
# Round-2 GPU call AC: bench lines of the other BASELINE configs with the final kernels (cfg3 bf16 decoder, cfg4 VOC
# shapes, cfg5 deterministic stress) at N=1, and ncu --set full of the cfg3 and cfg5-deterministic kernels.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
for wl in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl > "$out/bench_${wl}_r02ac.json" 2> "$out/bench_${wl}_r02ac.err"; echo "bench $wl exit $?" >> "$out/bench_${wl}_r02ac.err"
done
PROF="python bench.py --steps 1 --warmup 1 --layers 1 --regions 1 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 120 $PROF --workload cfg3 > "$out/prof_plain_cfg3_r02ac.log" 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 3 -f -o "$out/prof_cfg3_r02ac" $PROF --workload cfg3 > "$out/ncu_cfg3_r02ac.log" 2>&1
timeout 120 $PROF --workload cfg5 > "$out/prof_plain_cfg5_r02ac.log" 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_|det_" -c 12 -f -o "$out/prof_cfg5_r02ac" $PROF --workload cfg5 > "$out/ncu_cfg5_r02ac.log" 2>&1
for wl in cfg3 cfg4 cfg5; do python -c "
import json; d=json.loads(open('$out/bench_${wl}_r02ac.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$wl', d['value'], d['ms_per_step'], r['frac'], r.get('fwd_bwd_frac'), r['launch_ms'], r['fwd']['launch_ms'], d['e2e']['value'], d['clocks'])"; tail -1 "$out/bench_${wl}_r02ac.err"; done
tail -3 "$out/ncu_cfg3_r02ac.log"; tail -3 "$out/ncu_cfg5_r02ac.log"
