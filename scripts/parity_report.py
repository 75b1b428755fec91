#!/usr/bin/env python
"""Parity of the CUDA path against the oracle's fp64 arbiter on every test case: max-abs error, SNR,
and whether the index path (bypass mode) is bit-exact.  Writes one JSON object per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
import rocoder_b200 as rb
from cases import LARGE_CASES, SMALL_CASES

rows = []
for n, f, p, W in SMALL_CASES + LARGE_CASES:
    rng = np.random.default_rng(1)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    ref = ol.process_all(x, f, p, W, seed=1234, precision=64)
    ref32 = ol.process_all(x, f, p, W, seed=1234, precision=32)
    byp = ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1234) as e:
        got = e.process_all(x)
    with rb.Engine(2, 48000, W, f, 1.0, p, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
        gb = e.process_all(x)
    row = {"n_in": n, "factor": f, "pitch": p, "window": W, "out_len": int(ref.shape[1]),
           "cuda_vs_f64_max_abs": float(np.abs(got - ref).max()) if ref.size else 0.0,
           "cuda_vs_f64_snr_db": ol.snr_db(ref, got) if ref.any() else None,
           "oracle_f32_vs_f64_max_abs": float(np.abs(ref32 - ref).max()) if ref.size else 0.0,
           "peak": float(np.abs(ref).max()) if ref.size else 0.0,
           "index_path_bit_exact": bool(np.array_equal(gb, byp))}
    rows.append(row)
    print(json.dumps(row), flush=True)
worst = max(r["cuda_vs_f64_max_abs"] for r in rows)
snrs = [r["cuda_vs_f64_snr_db"] for r in rows if r["cuda_vs_f64_snr_db"] is not None]
print(json.dumps({"summary": {"cases": len(rows), "worst_max_abs": worst, "min_snr_db": min(snrs),
                              "all_index_paths_bit_exact": all(r["index_path_bit_exact"] for r in rows),
                              "tolerance": "1e-4 max-abs or SNR > 100 dB (BASELINE.json north_star)"}}))
