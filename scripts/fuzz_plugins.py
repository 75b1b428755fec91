#!/usr/bin/env python
"""Randomised parity sweep of the three device-side frequency-kernel routes against the oracle calling the
equivalent host kernel on the full spectrum: fused bin-local form (NVRTC builds the stretch kernel around it),
gather form (spectra through HBM), host C-ABI form.  usage: python scripts/fuzz_plugins.py [n_cases] [seed]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
import plugins
import rocoder_b200 as rb
from rocoder_b200 import _abi

LOCAL_TILT = """
float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 v) {
    float g = 0.25f + (float)bin / (float)n_bins;
    return make_float2((v.x * 0.6f - v.y * 0.8f) * g, (v.x * 0.8f + v.y * 0.6f) * g);
}
"""
GATHER_TILT = """
float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    float g = 0.25f + (float)bin / (float)n_bins;
    float2 v = spectrum[bin];
    return make_float2((v.x * 0.6f - v.y * 0.8f) * g, (v.x * 0.8f + v.y * 0.6f) * g);
}
"""
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad, done = [], 0
while done < n_cases:
    W = int(2 ** rng.integers(6, 17))
    p = int(rng.choice([1, 2, 3, 4, -2, -3, -4]))
    f = float(np.float32(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 8.0, rng.uniform(0.5, 10.0)])))
    c = ol.constants(f, p, W)
    step = c["sample_step_len"]
    if not (1 <= step <= W):
        continue
    n = int(rng.integers(1, W + int(rng.integers(1, 30)) * step + 1))
    if n * c["pitch_shifted_factor"] > 2e6:
        continue
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    seed = int(rng.integers(0, 2 ** 40))
    ref = ol.process_all(x, f, p, W, seed=seed, precision=64, plugin=plugins.tilt)
    res = {}
    try:
        with rb.Engine(2, 48000, W, f, 1.0, p, seed=seed) as e:
            e.push_local_device_frequency_kernel(LOCAL_TILT)
            res["fused"] = e.process_all(x)
        if done % 2 == 0:
            with rb.Engine(2, 48000, W, f, 1.0, p, seed=seed) as e:
                e.push_device_frequency_kernel(GATHER_TILT)
                res["gather"] = e.process_all(x)
        if done % 3 == 0:
            with rb.Engine(2, 48000, W, f, 1.0, p, seed=seed) as e:
                cb = _abi.APPLY_FN(plugins.tilt)
                _abi.check(_abi.lib().rcg_push_frequency_kernel(e._h, cb))
                res["host"] = e.process_all(x)
        for k, got in res.items():
            err = float(np.abs(got - ref).max())
            if got.shape != ref.shape or err > 1e-4 or (ref.any() and ol.snr_db(ref, got) <= 100.0):
                bad.append(dict(route=k, W=W, p=p, f=f, n=n, seed=seed, max_abs=err))
                print("FAIL", json.dumps(bad[-1]), flush=True)
    except Exception as ex:  # noqa
        bad.append(dict(W=W, p=p, f=f, n=n, seed=seed, error=repr(ex)[:300]))
        print("ERROR", json.dumps(bad[-1]), flush=True)
    done += 1
print(json.dumps({"cases": done, "failures": len(bad), "failed": bad[:20]}))
sys.exit(1 if bad else 0)
