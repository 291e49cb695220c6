#!/usr/bin/env bash
# Runs tools/debug/probe_lib.py step by step, each under a timeout; on a hang, attaches cuda-gdb for native stacks.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
log="$out/probe.log"
: > "$log"
env | grep -i "cuda\|nvidia" >> "$log" 2>&1
for lib in build/variants/lib_r01.so ir_ads_b200/libmsda_b200.so; do
  for step in book generic fast; do
    echo "=== $lib $step" >> "$log"
    python -u tools/debug/probe_lib.py "$lib" "$step" >> "$log" 2>&1 &
    pid=$!
    for t in $(seq 1 25); do sleep 1; kill -0 $pid 2>/dev/null || break; done
    if kill -0 $pid 2>/dev/null; then
      echo "--- still running after 25 s: native stacks" >> "$log"
      cat /proc/$pid/wchan >> "$log" 2>&1; echo >> "$log"
      timeout 60 /usr/local/cuda/bin/cuda-gdb-minimal -batch -ex "thread apply all bt 25" -p $pid >> "$log" 2>&1
      kill -9 $pid 2>/dev/null
    fi
    wait $pid 2>/dev/null
    echo "rc $?" >> "$log"
  done
done
nvidia-smi >> "$log" 2>&1
tail -150 "$log"
