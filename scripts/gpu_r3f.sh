#!/bin/bash
# W = 16384: spectral-step table entries from the Tensor Memory row (default build) against the global tables
# (variants_tmp/lib_spec_tables.so = the previous commit's library)
TAG=${TAG:-r3f}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 8 --pitch 2" "--factor 4 --pitch 1" "--factor 20 --pitch -3"; do
  for v in tables tmem tables tmem; do
    if [ $v = tables ]; then export RCG_LIB=variants_tmp/lib_spec_tables.so; else unset RCG_LIB; fi
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
tail -5 gpurun_out/${TAG}_err.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_plugin_path.py -m gpu -q -x 2>&1 | tail -4
