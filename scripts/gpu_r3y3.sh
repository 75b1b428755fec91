#!/bin/bash
# memcheck over the paths added after the last sanitizer run: window / envelope in Tensor Memory (W = 16384, full-rate
# modes), the rewritten lerp walk and decimating gather, the split load of the CTA pair with a negative pitch
mkdir -p gpurun_out
cat > /tmp/san_last.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import rocoder_b200 as rb
rng = np.random.default_rng(0)
for W, f, p, n in ((16384, 2.0, 1, 300000), (16384, 6.0, -3, 300000), (16384, 4.0, -2, 200000), (16384, 3.0, 3, 200000),
                   (4096, 5.0, -3, 100000), (1024, 3.0, 5, 50000), (65536, 6.0, -3, 500000), (16384, 4.0, 2, 200000)):
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1) as e:
        a = e.process_all(x)
    print("ok", W, f, p, a.shape, float(np.abs(a).sum()), flush=True)
PY
timeout 150 compute-sanitizer --tool memcheck python /tmp/san_last.py > gpurun_out/r3y_sanitize_memcheck_last.log 2>&1
grep -v "^=========     \|^========= $" gpurun_out/r3y_sanitize_memcheck_last.log | tail -n 12
