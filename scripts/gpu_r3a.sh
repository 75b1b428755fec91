#!/bin/bash
# round 2, session 2: factored constant-twiddle butterflies + one-product phase scaling, A/B on the same box.
# The (window, variant) libraries are prebuilt into variants_tmp/ (git-ignored; it travels with the snapshot):
#   base  -DRCG_BFLY_FACTORED=0 -DRCG_PHASE_TWO_PRODUCTS     bfly  -DRCG_BFLY_FACTORED=1 -DRCG_PHASE_TWO_PRODUCTS
#   phase -DRCG_BFLY_FACTORED=0                               both  -DRCG_BFLY_FACTORED=1
TAG=${TAG:-r3a}
WINDOWS=${WINDOWS:-"16384 4096 1024"}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for W in $WINDOWS; do
  for cfg in "--factor 8 --pitch 2" "--factor 4 --pitch 1" "--factor 20 --pitch -3"; do
    for v in base bfly phase both base both; do
      RCG_LIB=variants_tmp/lib_${v}_$W.so timeout 120 python scripts/kernel_time.py --tag $v --window $W $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log
    done
  done
done
TAG=$TAG python - <<'PY'
import json, os
for l in open("gpurun_out/%s_variants.jsonl" % os.environ["TAG"]):
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["tag"], d["W"], d["f"], d["p"], round(d["kernel_ms"], 4), round(d["min_ms"], 4), d["checksum"])
PY
