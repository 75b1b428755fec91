#!/bin/bash
# W = 65536: second sweep of the unroll factors of the load / overlap-add phases (variants_tmp/lib_pair_*.so)
TAG=${TAG:-r3r}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2" "--factor 20 --pitch -3"; do
  for v in base l2o4 l1o4 l2o8 l3o4 base l2o4 l1o4 l2o8 l3o4; do
    if [ $v = base ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_pair_$v.so; fi
    timeout 120 python scripts/kernel_time.py --tag $v --window 65536 $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
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
