#!/usr/bin/env bash
# Round-2 GPU call AH: validation of the library with the deterministic-path changes of calls AD-AG (whole GPU suite,
# smoke, default bench line) and an A/B of the NON-deterministic kernels against the previous library (their code did
# not change, but their parameter block did).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02ah.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02ah.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02ah.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02ah.log"
for v in head product head product; do
  echo "== $v" | tee -a "$out/sweep_r02ah.log"
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 python tools/sweep.py --iters 15 --workloads cfg2,cfg3,cfg5 2>&1 | grep -v "^\[" | cut -c1-130 | tee -a "$out/sweep_r02ah.log"
done
timeout 400 python bench.py > "$out/bench_r02ah.json" 2> "$out/bench_r02ah.err"; echo "bench exit $?" >> "$out/bench_r02ah.err"
tail -3 "$out/pytest_r02ah.log"; tail -2 "$out/smoke_r02ah.log"; python -c "
import json; d=json.load(open('$out/bench_r02ah.json')); print({k:d[k] for k in ('value','ms_per_step','timed_regions','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['fwd_bwd_frac'], d['e2e']['value'], d['cpu_baseline']['value'])"; tail -1 "$out/bench_r02ah.err"
