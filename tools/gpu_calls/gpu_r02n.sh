#!/usr/bin/env bash
# Round-2 GPU call N: the whole parity suite against a -DMSDA_DEBUG_BOUNDS build (every dereferenced corner offset
# and mask byte checked in-kernel; a violation prints and traps) -- the memcheck of the unclamped records.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
MSDA_B200_LIB=build/variants/lib_bounds.so timeout 1500 python -u -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread > "$out/pytest_bounds_r02n.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_bounds_r02n.log"
grep -c "msda bounds" "$out/pytest_bounds_r02n.log"; tail -6 "$out/pytest_bounds_r02n.log"
