#!/bin/bash
# W = 65536: the frame read once per CTA pair (split load, default build) against every CTA reading the whole frame
# (variants_tmp/lib_pair_whole.so), unroll factors on top of the split load; long-window / multi-device tests;
# racecheck of the pair kernel (the posted stores into the peer's buffer)
TAG=${TAG:-r3s}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_variants.jsonl
for cfg in "--factor 4 --pitch 1" "--factor 8 --pitch 2" "--factor 20 --pitch -3"; do
  for v in split whole l4o8 l1o8 l2o16 l8o8 split whole l4o8 l1o8 l2o16 l8o8; do
    if [ $v = split ]; then unset RCG_LIB; else export RCG_LIB=variants_tmp/lib_pair_$v.so; fi
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
( time timeout 900 python -m pytest tests/test_launch_geometry.py tests/test_any_window.py tests/test_full_size_configs.py tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 8 gpurun_out/${TAG}_pytest_gpu.log
cat > /tmp/san_pair.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import rocoder_b200 as rb
rng = np.random.default_rng(0)
for W, f, p in ((65536, 2.0, 1), (65536, 4.0, 2), (65536, 3.0, -2)):
    xx = ((rng.random((2, 400000), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1) as e:
        a = e.process_all(xx)
    print("ok", W, f, p, float(np.abs(a).sum()), flush=True)
PY
timeout 600 compute-sanitizer --tool racecheck python /tmp/san_pair.py > gpurun_out/${TAG}_sanitize_pair_racecheck.log 2>&1; tail -n 3 gpurun_out/${TAG}_sanitize_pair_racecheck.log
timeout 300 compute-sanitizer --tool memcheck python /tmp/san_pair.py > gpurun_out/${TAG}_sanitize_pair_memcheck.log 2>&1; tail -n 3 gpurun_out/${TAG}_sanitize_pair_memcheck.log
