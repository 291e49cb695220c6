#!/usr/bin/env bash
# Round-2 GPU call AI: the experiment-build tests against the rebuilt experiment library (build/variants/lib_exp.so,
# same sources as the product) and the cfg5 bench line with the final deterministic path.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
MSDA_B200_LIB=build/variants/lib_exp.so timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --timeout 200 --timeout-method=thread -k "fold or row_orders or pathological or determin" > "$out/pytest_exp_r02ai.log" 2>&1; echo "pytest exit $?" >> "$out/pytest_exp_r02ai.log"
timeout 300 python bench.py --workload cfg5 > "$out/bench_cfg5_r02ai.json" 2> "$out/bench_cfg5_r02ai.err"; echo "bench exit $?" >> "$out/bench_cfg5_r02ai.err"
tail -3 "$out/pytest_exp_r02ai.log"; python -c "
import json; d=json.loads(open('$out/bench_cfg5_r02ai.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['fwd_bwd_frac'], r['launch_ms'], r['fwd']['launch_ms'], d['gpu_launches'])"; tail -1 "$out/bench_cfg5_r02ai.err"
