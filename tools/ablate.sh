#!/usr/bin/env bash
# Builds ablation variants of the library into build/variants/ (never the product .so; results are WRONG on purpose):
#   no_red     backward without the grad_value scatter   -> cost of the gather + reductions alone
#   no_gather  backward without the value gather          -> cost of the scatter alone
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
bash "${here}/build_variant.sh" no_red -DMSDA_EXP_NO_RED &
bash "${here}/build_variant.sh" no_gather -DMSDA_EXP_NO_GATHER &
wait
ls -la "${here}/../build/variants/"
