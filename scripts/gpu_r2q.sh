#!/bin/bash
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 1500 python scripts/fuzz_parity.py 300 11 > gpurun_out/${TAG}_fuzz_parity.log 2>&1; tail -3 gpurun_out/${TAG}_fuzz_parity.log | cut -c1-600
timeout 600 python scripts/fuzz_plugins.py 40 7 > gpurun_out/${TAG}_fuzz_plugins.log 2>&1; tail -2 gpurun_out/${TAG}_fuzz_plugins.log | cut -c1-400
cat > /tmp/san_extra.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import rocoder_b200 as rb
rng = np.random.default_rng(0)
x = ((rng.random((2, 60000), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
for W, f, p in ((1000, 2.0, 1), (600, 3.0, 2), (1500, 4.0, -3)):
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1) as e:
        a = e.process_all(x)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1, devices=[0, 0, 0], flags=rb.RCG_FLAG_SHARD_ALWAYS) as e:
        assert np.array_equal(e.process_all(x), a)
    print("ok chirp-z", W, f, p, flush=True)
for W, f, p in ((16384, 8.0, 2), (4096, 20.0, -3), (65536, 2.0, 1)):
    xx = ((rng.random((2, 900000), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1) as e:
        a = e.process_all(xx)
    for fl in (rb.RCG_FLAG_SHARD_ALWAYS, rb.RCG_FLAG_SHARD_ALWAYS | rb.RCG_FLAG_SHARD_HALO):
        with rb.Engine(2, 48000, W, f, 1.0, p, seed=1, devices=[0, 0, 0], flags=fl) as e:
            assert np.array_equal(e.process_all(xx), a)
    print("ok devices", W, f, p, flush=True)
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python /tmp/san_extra.py > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1; tail -2 gpurun_out/${TAG}_sanitize_${tool}.log
done
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/${TAG}_sanitize_small_memcheck.log 2>&1; tail -2 gpurun_out/${TAG}_sanitize_small_memcheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_b_ncu.log 2>&1
grep -c "stretch_kernel" gpurun_out/${TAG}_launches.csv
