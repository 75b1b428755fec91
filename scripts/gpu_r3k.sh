#!/bin/bash
# 16-column Tensor Memory loads (default build) against 8-column ones (variants_tmp/lib_chunk8.so) at the other window
# sizes; then the whole GPU suite on the default build
TAG=${TAG:-r3k}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for W in 1024 4096 8192 32768 65536; do
  for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2"; do
    for v in x8 x16 x8 x16; do
      if [ $v = x8 ]; then export RCG_LIB=variants_tmp/lib_chunk8.so; else unset RCG_LIB; fi
      timeout 120 python scripts/kernel_time.py --tag $v --window $W $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
    done
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
