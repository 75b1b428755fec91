#!/bin/bash
# phase words -> phases with a funnel shift and one fused multiply-add (default build) against shift / int-to-float /
# multiply (variants_tmp/lib_before_phase.so): bit-identical output expected (checksums), timing at W = 1024 .. 65536
TAG=${TAG:-r3w}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--window 16384 --factor 8 --pitch 2" "--window 16384 --factor 4 --pitch 1" "--window 1024 --factor 4 --pitch 1" "--window 4096 --factor 4 --pitch 1" "--window 65536 --factor 4 --pitch 1" "--window 16384 --factor 20 --pitch -3"; do
  for v in before fused before fused; do
    if [ $v = fused ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_before_phase.so; fi
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
