#!/bin/bash
# W = 16384, full-rate modes: window and envelope from the Tensor Memory row (default build) against global tables
# (variants_tmp/lib_before_wintmem.so); then the whole GPU suite on the default build
TAG=${TAG:-r3x}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--window 16384 --factor 4 --pitch 1" "--window 16384 --factor 20 --pitch -3" "--window 16384 --factor 2 --pitch 1 --channels 8 --seconds 120" "--window 16384 --factor 8 --pitch -2" "--window 16384 --factor 6 --pitch 3" "--window 16384 --factor 8 --pitch 2"; do
  for v in before wintmem before wintmem; do
    if [ $v = wintmem ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_before_wintmem.so; fi
    timeout 120 python scripts/kernel_time.py --tag $v $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
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
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 12 gpurun_out/${TAG}_pytest_gpu.log
