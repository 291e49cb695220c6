#!/usr/bin/env bash
# Round-2 GPU call W: last check of the committed tree -- whole GPU suite, smoke, the default bench line.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02w.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02w.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02w.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02w.log"
timeout 400 python bench.py > "$out/bench_r02w.json" 2> "$out/bench_r02w.err"; echo "bench exit $?" >> "$out/bench_r02w.err"
tail -3 "$out/pytest_r02w.log"; tail -2 "$out/smoke_r02w.log"; python -c "
import json; d=json.load(open('$out/bench_r02w.json')); print({k:d[k] for k in ('value','ms_per_step','timed_regions','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['fwd_bwd_frac'], d['e2e']['value'], d['cpu_baseline']['value'])"; tail -1 "$out/bench_r02w.err"
