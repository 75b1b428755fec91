#!/bin/bash
# round 2: which phases gain from packed FP32?  variants of the W = 16384 kernels, three workloads each
mkdir -p gpurun_out/variants
OUT=gpurun_out/r2e_variants.jsonl; : > $OUT
for spec in "pk1111:-DRCG_PK_DFT=1 -DRCG_PK_TW=1 -DRCG_PK_SPEC=1 -DRCG_PK_IO=1" \
            "pk0000:-DRCG_PK_DFT=0 -DRCG_PK_TW=0 -DRCG_PK_SPEC=0 -DRCG_PK_IO=0" \
            "pk1101:-DRCG_PK_DFT=1 -DRCG_PK_TW=1 -DRCG_PK_SPEC=0 -DRCG_PK_IO=1" \
            "pk1001:-DRCG_PK_DFT=1 -DRCG_PK_TW=0 -DRCG_PK_SPEC=0 -DRCG_PK_IO=1" \
            "pk0100:-DRCG_PK_DFT=0 -DRCG_PK_TW=1 -DRCG_PK_SPEC=0 -DRCG_PK_IO=0" \
            "pk1000:-DRCG_PK_DFT=1 -DRCG_PK_TW=0 -DRCG_PK_SPEC=0 -DRCG_PK_IO=0" \
            "pk0010:-DRCG_PK_DFT=0 -DRCG_PK_TW=0 -DRCG_PK_SPEC=1 -DRCG_PK_IO=0"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  lib=gpurun_out/variants/lib_${name}.so
  python - "$lib" 13 $flags <<'PY' 2> gpurun_out/variants/${name}.build.log || { echo "{\"tag\": \"$name\", \"error\": \"build failed\"}" >> $OUT; continue; }
import sys
sys.path.insert(0, ".")
from rocoder_b200 import build
build.build_product(force=True, extra_flags=tuple(sys.argv[3:]), out=sys.argv[1], only_log2h=int(sys.argv[2]))
PY
  RCG_LIB=$lib python scripts/kernel_time.py --tag "${name}_cfg2" >> $OUT 2>> gpurun_out/variants/${name}.build.log
  RCG_LIB=$lib python scripts/kernel_time.py --tag "${name}_f4p1" --factor 4 --pitch 1 >> $OUT 2>> gpurun_out/variants/${name}.build.log
  RCG_LIB=$lib python scripts/kernel_time.py --tag "${name}_f20pm3" --factor 20 --pitch -3 >> $OUT 2>> gpurun_out/variants/${name}.build.log
  rm -rf $lib ${lib}.objs
done
python - <<'PY'
import json
for l in open("gpurun_out/r2e_variants.jsonl"):
    try:
        d = json.loads(l); print(d["tag"], d.get("kernel_ms"), d.get("error", ""))
    except Exception: print(l.strip()[:200])
PY
