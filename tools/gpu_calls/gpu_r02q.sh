#!/usr/bin/env bash
# Round-2 GPU call Q: the FINAL library of the round -- whole parity suite, experiment tests,
# smoke, bench line + reference arm, launch list + ncu --set full of the two hot kernels, sweep of every config.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q -s --timeout 300 --timeout-method=thread > "$out/pytest_r02q.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02q.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological" > "$out/pytest_exp_r02q.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02q.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02q.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02q.log"
timeout 400 python bench.py > "$out/bench_r02q.json" 2> "$out/bench_r02q.err"; echo "bench exit $?" >> "$out/bench_r02q.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref_r02q.json" 2> "$out/bench_ref_r02q.err"
timeout 200 python tools/sweep.py --iters 15 --workloads cfg2,cfg2_bf16,cfg3,cfg3_f32,cfg4,cfg5 > "$out/sweep_r02q.log" 2>&1
timeout 200 python tools/sweep.py --iters 10 --det --workloads cfg2,cfg5 >> "$out/sweep_r02q.log" 2>&1
PROF="python bench.py --steps 1 --warmup 3 --layers 1 --regions 1 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 120 $PROF > "$out/prof_plain_r02q.log" 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_r02q.csv" $PROF > "$out/ncu_launches_r02q.log" 2>&1
timeout 120 $PROF > "$out/prof_plain2_r02q.log" 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o "$out/prof_r02q" $PROF > "$out/ncu_full_r02q.log" 2>&1
tail -4 "$out/pytest_r02q.log"; tail -3 "$out/pytest_exp_r02q.log"; tail -2 "$out/smoke_r02q.log"; python -c "
import json; d=json.load(open('$out/bench_r02q.json')); print({k:d[k] for k in ('value','ms_per_step','timed_regions','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['fwd_bwd_frac'], d['roofline']['launch_ms'], d['roofline']['fwd']['launch_ms'], d['e2e']['value'])"; cat "$out/sweep_r02q.log"; tail -3 "$out/ncu_full_r02q.log"
