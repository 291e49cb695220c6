#!/usr/bin/env bash
# Round-2 GPU call Y: product library reduced to one row order per pass (LINEAR forward, STRIP backward) -- whole suite on the
# product, order / fold tests on the experiment build, smoke, default bench line.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > "$out/pytest_r02y.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02y.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological or random_problem" > "$out/pytest_exp_r02y.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02y.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02y.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02y.log"
timeout 400 python bench.py --no-cpu-baseline > "$out/bench_r02y.json" 2> "$out/bench_r02y.err"; echo "bench exit $?" >> "$out/bench_r02y.err"
tail -3 "$out/pytest_r02y.log"; tail -3 "$out/pytest_exp_r02y.log"; tail -2 "$out/smoke_r02y.log"; python -c "
import json; d=json.load(open('$out/bench_r02y.json')); print({k:d[k] for k in ('value','ms_per_step','timed_regions','gpu_launches')}); print(d['roofline']['frac'], d['roofline']['fwd_bwd_frac'], d['e2e']['value'])"
