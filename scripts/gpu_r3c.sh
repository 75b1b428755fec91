#!/bin/bash
# pass twiddles in Tensor Memory (default build) against the shared-memory tables (-DRCG_TW_TMEM=0), then parity
TAG=${TAG:-r3c}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 8 --pitch 2" "--factor 4 --pitch 1" "--factor 20 --pitch -3" "--factor 2 --pitch 3"; do
  for v in smem tmem smem tmem; do
    if [ $v = smem ]; then export RCG_LIB=variants_tmp/lib_smemtw_16384.so; else unset RCG_LIB; fi
    timeout 120 python scripts/kernel_time.py --tag $v --window 16384 $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log
  done
done
unset RCG_LIB
TAG=$TAG python - <<'PY'
import json, os
for l in open("gpurun_out/%s_variants.jsonl" % os.environ["TAG"]):
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["tag"], d["W"], d["f"], d["p"], round(d["kernel_ms"], 4), round(d["min_ms"], 4), d["checksum"])
PY
tail -5 gpurun_out/${TAG}_err.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
