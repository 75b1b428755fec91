#!/bin/bash
# W = 1024 / 4096 / 8192: Tensor Memory twiddles + spectral entries (default build) against the global tables
# (variants_tmp/lib_small_tables_<W>.so, -DRCG_TW_TMEM_SMALL=0); then the whole GPU suite
TAG=${TAG:-r3g}
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
unset RCG_LIB
TAG=$TAG python - <<'PY'
import json, os
for l in open("gpurun_out/%s_variants.jsonl" % os.environ["TAG"]):
    try: d = json.loads(l); print(d["tag"], d["W"], d["f"], d["p"], round(d["kernel_ms"], 4), round(d["min_ms"], 4), d["checksum"])
    except Exception: print(l[:200])
PY
tail -5 gpurun_out/${TAG}_err.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
