#!/bin/bash
# W = 32768 (512 threads per frame): unroll factors of the load / overlap-add loops (variants_tmp/lib_w32k_*.so);
# ncu capture of the W = 65536 pair kernel as it ships
TAG=${TAG:-r3t}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2"; do
  for v in base u22 u24 u28 u48 u82 base u22 u24 u28 u48 u82; do
    if [ $v = base ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_w32k_$v.so; fi
    timeout 120 python scripts/kernel_time.py --tag $v --window 32768 $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_pair -s 2 -c 1 -o gpurun_out/${TAG}_pair python scripts/kernel_time.py --tag pair --window 65536 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_pair.log 2>&1
tail -n 2 gpurun_out/${TAG}_pair.log
