#!/usr/bin/env bash
# Round-2 GPU call D: new tests (epilogue), experiment-build tests, N=1 column of the scaling table.
set -u
out=gpurun_out
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 300 python -u -m pytest tests/test_epilogue_gpu.py tests/test_encoder.py tests/test_module_gpu.py -m gpu -x -q --timeout 120 --timeout-method=thread > "$out/pytest_r02d.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_r02d.log"
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological" > "$out/pytest_exp_r02d.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02d.log"
bash tools/scaling_r02.sh 1 > "$out/scale_r02_n1.txt" 2>&1
timeout 200 python tools/bench_layer.py > "$out/bench_layer_r02d.json" 2> "$out/bench_layer_r02d.err"
tail -4 "$out/pytest_r02d.log"; tail -4 "$out/pytest_exp_r02d.log"; cat "$out/scale_r02_n1.txt" | head -12; cat "$out/bench_layer_r02d.json"; tail -3 "$out/bench_layer_r02d.err"
