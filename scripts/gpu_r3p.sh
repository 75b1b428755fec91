#!/bin/bash
# W = 65536: unroll factors of the load and overlap-add phases after the LDS fix (variants_tmp/lib_pair_*.so built on
# the CPU box), an ncu capture of the pair kernel, the long-window / multi-device tests
TAG=${TAG:-r3p}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2"; do
  for v in base load2 load8 ola1 ola4 base load2 load8 ola1 ola4; do
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stretch_pair -s 2 -c 1 -o gpurun_out/${TAG}_pair python scripts/kernel_time.py --tag pair --window 65536 --factor 4 --pitch 1 --reps 1 > gpurun_out/${TAG}_pair.log 2>&1
tail -n 2 gpurun_out/${TAG}_pair.log
( time timeout 900 python -m pytest tests/test_launch_geometry.py tests/test_any_window.py tests/test_full_size_configs.py tests/test_multi_device.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 8 gpurun_out/${TAG}_pytest_gpu.log
