#!/usr/bin/env bash
# Round-2 GPU call G: verification of the final library -- whole parity suite (product build), experiment tests
# (experiment build), smoke, the bench line and the reference arm.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q -s --timeout 300 --timeout-method=thread > "$out/pytest_r02g.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02g.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological" > "$out/pytest_exp_r02g.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02g.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02g.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02g.log"
timeout 400 python bench.py > "$out/bench_r02g.json" 2> "$out/bench_r02g.err"; echo "bench exit $?" >> "$out/bench_r02g.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref_r02g.json" 2> "$out/bench_ref_r02g.err"
timeout 200 python tools/sweep.py --iters 15 --workloads cfg2,cfg2_bf16,cfg3,cfg3_f32,cfg4,cfg5 > "$out/sweep_r02g.log" 2>&1
timeout 200 python tools/sweep.py --iters 10 --det --workloads cfg2,cfg5 >> "$out/sweep_r02g.log" 2>&1
grep -c "full-size parity" "$out/pytest_r02g.log"; tail -3 "$out/pytest_r02g.log"; tail -3 "$out/pytest_exp_r02g.log"; tail -2 "$out/smoke_r02g.log"; python -c "
import json; d=json.load(open('$out/bench_r02g.json')); print({k:d[k] for k in ('value','ms_per_step','timed_regions','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['fwd_bwd_frac'], d['roofline']['launch_ms'], d['roofline']['fwd']['launch_ms'], d['e2e']['value'])"; cat "$out/sweep_r02g.log"
