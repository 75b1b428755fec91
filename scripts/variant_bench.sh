#!/bin/bash
# Build variants of the kernels (one window size) with different -D flags on the GPU box and time them.
# env: ONLY_LOG2H (log2(W)-1 to compile, default 13), KT_ARGS (arguments of kernel_time.py)
# usage: bash scripts/variant_bench.sh "name1:-DFOO=1 -DBAR=2" "name2:..." ...
mkdir -p gpurun_out/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  out=gpurun_out/variants/lib_${name}.so
  python - "$out" ${ONLY_LOG2H:-13} $flags <<'PY' 2> gpurun_out/variants/${name}.build.log || { echo "{\"tag\": \"$name\", \"error\": \"build failed\"}"; continue; }
import sys
sys.path.insert(0, ".")
from rocoder_b200 import build
build.build_product(force=True, extra_flags=tuple(sys.argv[3:]), out=sys.argv[1], only_log2h=int(sys.argv[2]))
PY
  RCG_LIB=$out python scripts/kernel_time.py --tag "$name" $KT_ARGS 2>> gpurun_out/variants/${name}.build.log
done
