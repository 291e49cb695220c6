#!/usr/bin/env bash
# Round-2 GPU call AF: per-kernel durations of the deterministic backward at cfg 5 (ncu launch lists) for the previous
# library (head), the multiply-high-division variant (fd) and the candidate product.
set -u
out=gpurun_out; mkdir -p "$out"; export PYTHONUNBUFFERED=1
for v in head fd product; do
  lib="build/variants/lib_${v}.so"; [[ $v == product ]] && lib="ir_ads_b200/libmsda_b200.so"
  MSDA_B200_LIB="$lib" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"msda_|det_" -s 26 -c 13 --csv --log-file "$out/launches_det_${v}_r02af.csv" python tools/sweep.py --iters 1 --det --workloads cfg5 > "$out/ncu_det_${v}_r02af.log" 2>&1
  echo "== $v"; python - "$out/launches_det_${v}_r02af.csv" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:]:
    print(f"{float(r[vi].replace(',', '')) / 1e3:9.1f} us  {r[ki][:90]}")
PY
done
