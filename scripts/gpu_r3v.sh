#!/bin/bash
# negative pitch multiples: the lerp walk with scalar shared-memory loads, running pointers and a fixed trip count
# (default build) against the build before it (variants_tmp/lib_before_lerp.so)
TAG=${TAG:-r3v}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--window 16384 --factor 20 --pitch -3" "--window 16384 --factor 8 --pitch -2" "--window 4096 --factor 20 --pitch -3" "--window 1024 --factor 20 --pitch -3" "--window 65536 --factor 20 --pitch -3" "--window 16384 --factor 12 --pitch -5"; do
  for v in before lerp before lerp; do
    if [ $v = lerp ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_before_lerp.so; fi
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
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_any_window.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 5 gpurun_out/${TAG}_pytest_gpu.log
