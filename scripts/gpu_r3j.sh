#!/bin/bash
# W = 1024 / 4096 / 8192 with the occupancy fix (three 128-column CTAs per SM): Tensor Memory build against the global
# tables (variants_tmp/lib_small_tables_<W>.so); W = 16384: 16-column Tensor Memory loads against 8-column ones
TAG=${TAG:-r3j}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for W in 1024 4096 8192; do
  for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2" "--factor 20 --pitch -3"; do
    for v in tables tmem tables tmem; do
      if [ $v = tables ]; then export RCG_LIB=variants_tmp/lib_small_tables_$W.so; else unset RCG_LIB; fi
      timeout 120 python scripts/kernel_time.py --tag $v --window $W $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
    done
  done
done
for cfg in "--factor 8 --pitch 2" "--factor 4 --pitch 1" "--factor 20 --pitch -3"; do
  for v in x8 x16 x8 x16; do
    if [ $v = x16 ]; then export RCG_LIB=variants_tmp/lib_chunk16_16384.so; else unset RCG_LIB; fi
    timeout 120 python scripts/kernel_time.py --tag $v --window 16384 $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
  done
done
unset RCG_LIB
TAG=$TAG python - <<'PY'
import json, os
for l in open("gpurun_out/%s_variants.jsonl" % os.environ["TAG"]):
    try: d = json.loads(l); print(d["tag"], d["W"], d["f"], d["p"], round(d["kernel_ms"], 4), round(d["min_ms"], 4), d["checksum"])
    except Exception: print(l[:200])
PY
tail -n 5 gpurun_out/${TAG}_err.log
