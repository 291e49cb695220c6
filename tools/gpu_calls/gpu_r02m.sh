#!/usr/bin/env bash
# Round-2 GPU call M: compute-sanitizer memcheck over the parity tests that exercise padded / gated / masked corners
# (compute-sanitizer turned out to be closed on this GPU pool; tools/gpu_calls/gpu_r02n.sh -- the -DMSDA_DEBUG_BOUNDS build -- replaced it)
# (the records hold UNCLAMPED offsets now: prove that no corner outside the map is ever dereferenced).
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
SAN="compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20"
timeout 1500 $SAN python -u -m pytest tests/test_parity_gpu.py tests/test_module_gpu.py tests/test_dcnv3.py -m gpu -x -q --timeout 900 --timeout-method=thread \
  -k "golden or channels_against or fast_and_generic or row_orders or random_problem or empty_and_ragged or nan_and_inf or non_finite or bookkeeping or padding_mask or discards or deterministic_sorted or skips_the_scatter or dcnv3" \
  > "$out/memcheck_r02m.log" 2>&1; echo "memcheck exit $?" >> "$out/memcheck_r02m.log"
grep -c "Invalid\|out of bounds" "$out/memcheck_r02m.log"; tail -8 "$out/memcheck_r02m.log"
