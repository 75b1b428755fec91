#!/bin/bash
# Tensor Memory twiddles: next chunk issued before the last one is used (default build) against load + wait in one
# statement (variants_tmp/lib_tmem_nopipe.so); ncu capture of the headline kernel
TAG=${TAG:-r3d}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 8 --pitch 2" "--factor 4 --pitch 1" "--factor 20 --pitch -3"; do
  for v in nopipe pipe nopipe pipe; do
    if [ $v = nopipe ]; then export RCG_LIB=variants_tmp/lib_tmem_nopipe.so; else unset RCG_LIB; fi
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stretch_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_b_ncu2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
