#!/bin/bash
# W = 65536 after the shared-memory pointers of the CTA pair kept their address space (LDS / STS instead of generic
# LD / ST); compute-sanitizer synccheck on a kernel that does nothing but tcgen05.alloc / st / ld / dealloc; the
# long-window and geometry tests
TAG=${TAG:-r3n}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2" "--factor 20 --pitch -3" "--factor 2 --pitch 1 --channels 8"; do
  for v in lds lds; do
    timeout 120 python scripts/kernel_time.py --tag $v --window 65536 $cfg >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log || echo "{\"tag\": \"$v\", \"error\": \"rc $?\"}" >> gpurun_out/${TAG}_variants.jsonl
  done
done
timeout 120 python scripts/kernel_time.py --tag sanity --window 32768 --factor 4 --pitch 1 >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log
timeout 120 python scripts/kernel_time.py --tag sanity --window 16384 --factor 8 --pitch 2 >> gpurun_out/${TAG}_variants.jsonl 2>> gpurun_out/${TAG}_err.log
TAG=$TAG python - <<'PY'
import json, os
for l in open("gpurun_out/%s_variants.jsonl" % os.environ["TAG"]):
    try: d = json.loads(l); print(d["tag"], d["W"], d["f"], d["p"], round(d["kernel_ms"], 4), round(d["min_ms"], 4), d["checksum"], d.get("geometry"))
    except Exception: print(l[:200])
PY
tail -n 5 gpurun_out/${TAG}_err.log
( cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu ) 2>&1 | tail -n 3
timeout 300 compute-sanitizer --tool synccheck scripts/microbench/tmem_ld > gpurun_out/${TAG}_synccheck_tmem_microbench.log 2>&1
grep "^========= [A-Z]" gpurun_out/${TAG}_synccheck_tmem_microbench.log | sort | uniq -c | head -n 8
( time timeout 900 python -m pytest tests/test_launch_geometry.py tests/test_any_window.py tests/test_full_size_configs.py tests/test_multi_device.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 8 gpurun_out/${TAG}_pytest_gpu.log
