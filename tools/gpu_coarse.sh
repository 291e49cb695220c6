#!/usr/bin/env bash
# One gpurun call for the coarse-level shared-memory accumulation: its parity tests, an A/B sweep
# (flags 0 = default concurrent, 128 = off, 512 = serial), occupancy experiments, the per-kernel ncu launch list.
set -u
out=gpurun_out
mkdir -p "$out"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "coarse" > "$out/pytest_coarse.log" 2>&1; echo "exit $?" >> "$out/pytest_coarse.log"
tail -15 "$out/pytest_coarse.log"
timeout 600 python tools/sweep.py --workloads cfg2 --dists model,test --flags 0,128,512 --iters 10 > "$out/sweep_coarse_cfg2.jsonl" 2>&1
cat "$out/sweep_coarse_cfg2.jsonl"
echo "--- main kernel alone (coarse kernel skipped), 4 / 3 / 2 CTAs per SM"
export MSDA_B200_LIB=build/variants/lib_knobs.so   # tools/ablate.sh builds it
for pad in 0 43000 63000; do
  MSDA_EXP_SKIP_COARSE_KERNEL=1 MSDA_EXP_BWD_SMEM_PAD=$pad timeout 300 python tools/sweep.py --workloads cfg2 --flags 512 --iters 10 2>&1 | tail -1
done
echo "--- all reds, 3 / 2 CTAs per SM"
for pad in 43000 63000; do
  MSDA_EXP_BWD_SMEM_PAD=$pad timeout 300 python tools/sweep.py --workloads cfg2 --flags 128 --iters 10 2>&1 | tail -1
done
timeout 600 python tools/sweep.py --workloads cfg3,cfg3_f32,cfg5,cfg2_bf16,cfg4 --flags 0,128 --iters 10 > "$out/sweep_coarse_other.jsonl" 2>&1
cat "$out/sweep_coarse_other.jsonl"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda_ -c 30 --csv --log-file "$out/launches_coarse.csv" \
  python tools/sweep.py --workloads cfg2 --flags 512 --iters 2 > "$out/ncu_coarse.log" 2>&1
grep -E "msda_" "$out/launches_coarse.csv" | awk -F'","' '{print substr($5,1,70), $NF}' | tail -9
