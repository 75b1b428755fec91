#!/bin/bash
# packed FP32 at the small windows (issue-limited there): W = 1024 and 4096, scalar vs packed
mkdir -p gpurun_out/variants
OUT=gpurun_out/r2t_small_window_variants.jsonl; : > $OUT
for L in 9 11 7; do
W=$((1 << (L + 1)))
for spec in "scalar:-DRCG_PACKED_FP32=0" "packed:-DRCG_PACKED_FP32=1" "pk_dft_only:-DRCG_PK_DFT=1" "pk_dft_tw:-DRCG_PK_DFT=1 -DRCG_PK_TW=1"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  lib=gpurun_out/variants/lib_${name}_${L}.so
  python - "$lib" $L $flags <<'PY' 2> gpurun_out/variants/${name}_${L}.build.log || { echo "{\"tag\": \"$name $W\", \"error\": \"build failed\"}" >> $OUT; continue; }
import sys
sys.path.insert(0, ".")
from rocoder_b200 import build
build.build_product(force=True, extra_flags=tuple(sys.argv[3:]), out=sys.argv[1], only_log2h=int(sys.argv[2]))
PY
  RCG_LIB=$lib python scripts/kernel_time.py --tag "${name}_w${W}_f4p1" --window $W --factor 4 --pitch 1 >> $OUT 2>> gpurun_out/variants/${name}_${L}.build.log
  RCG_LIB=$lib python scripts/kernel_time.py --tag "${name}_w${W}_f8p2" --window $W --factor 8 --pitch 2 >> $OUT 2>> gpurun_out/variants/${name}_${L}.build.log
  rm -rf $lib ${lib}.objs
done
done
python - <<'PY'
import json
for l in open("gpurun_out/r2t_small_window_variants.jsonl"):
    try:
        d = json.loads(l); print(d["tag"], d.get("kernel_ms"), d.get("error", ""))
    except Exception: print(l.strip()[:200])
PY
