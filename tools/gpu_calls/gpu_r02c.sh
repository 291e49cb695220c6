#!/usr/bin/env bash
# Round-2 GPU call C: the whole parity suite on the product library, the experiment tests on the experiment build,
# the bench line, launch list + ncu captures, a single-GPU run of the training-step mode.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 1200 python -u -m pytest tests -m gpu -x -q -s --timeout 300 --timeout-method=thread > "$out/pytest_r02c.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02c.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 300 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 120 --timeout-method=thread -k "fold or row_orders or pathological" > "$out/pytest_exp_r02c.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02c.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_r02c.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_r02c.log"
timeout 400 python bench.py --steps 10 --warmup 3 > "$out/bench_r02c.json" 2> "$out/bench_r02c.err"; echo "bench exit $?" >> "$out/bench_r02c.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref_r02c.json" 2> "$out/bench_ref_r02c.err"
for cfg in "--workload cfg4 --mode train --scaling strong --total-batch 16" "--workload cfg4 --mode train --scaling strong --total-batch 16 --no-graph" "--workload cfg4 --mode train --scaling strong --total-batch 2 --amp"; do
  timeout 200 python bench.py $cfg --steps 10 --warmup 3 >> "$out/bench_train_r02c.jsonl" 2>> "$out/bench_train_r02c.err"
done
PROF="python bench.py --steps 1 --warmup 3 --layers 1 --regions 1 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 120 $PROF > "$out/prof_plain_r02c.log" 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_r02c.csv" $PROF > "$out/ncu_launches_r02c.log" 2>&1
timeout 120 $PROF > "$out/prof_plain2_r02c.log" 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o "$out/prof_r02c" $PROF > "$out/ncu_full_r02c.log" 2>&1
tail -6 "$out/pytest_r02c.log"; tail -4 "$out/pytest_exp_r02c.log"; tail -3 "$out/smoke_r02c.log"; cat "$out/bench_r02c.json"; tail -3 "$out/bench_r02c.err"; cat "$out/bench_train_r02c.jsonl"; tail -5 "$out/bench_train_r02c.err"
