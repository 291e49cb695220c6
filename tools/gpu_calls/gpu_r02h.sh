#!/usr/bin/env bash
# Round-2 GPU call H: bf16 backward through an L2-resident, image-chunked float scratch (tests + A/B timing)
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_module_gpu.py -m gpu -x -q -s --timeout 200 --timeout-method=thread -k "bf16 or fused or cfg3 or random_problem or autocast" > "$out/pytest_r02h.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02h.log"
{
echo "== bf16 backward: chunked L2-resident scratch (flags 0) vs full-batch scratch (flags 32768)"
timeout 300 python -u tools/sweep.py --iters 20 --workloads cfg3,cfg2_bf16 --flags 0,32768
} > "$out/sweep_bf16_r02h.log" 2>&1
timeout 200 python tools/bench_module.py > "$out/bench_module_r02h.json" 2> "$out/bench_module_r02h.err"
tail -3 "$out/pytest_r02h.log"; cat "$out/sweep_bf16_r02h.log"; tail -c 1500 "$out/bench_module_r02h.json"
