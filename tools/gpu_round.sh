#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, a short bench, then the ncu launch list and one
# --set full capture of the fast kernels (each only after the same command exited 0 without ncu).
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
set -u
tag="${1:-r01}"
out=gpurun_out
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu_${tag}.csv" 2>&1
python -m pytest tests -m gpu -x -q > "$out/pytest_${tag}.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_${tag}.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke_${tag}.log" 2>&1; echo "smoke exit $?" >> "$out/smoke_${tag}.log"
python bench.py --steps 10 --warmup 3 > "$out/bench_${tag}.json" 2> "$out/bench_${tag}.err"; echo "bench exit $?" >> "$out/bench_${tag}.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref_${tag}.json" 2> "$out/bench_ref_${tag}.err"
PROF="python bench.py --steps 1 --warmup 3 --layers 1 --no-cpu-baseline --no-e2e --no-ref-cuda"
$PROF > "$out/prof_plain_${tag}.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_${tag}.csv" $PROF > "$out/ncu_launches_${tag}.log" 2>&1
$PROF > "$out/prof_plain2_${tag}.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o "$out/prof_${tag}" $PROF > "$out/ncu_full_${tag}.log" 2>&1
ls -la "$out"
tail -5 "$out/pytest_${tag}.log"; cat "$out/smoke_${tag}.log" | tail -3; cat "$out/bench_${tag}.json"; tail -3 "$out/bench_${tag}.err"
