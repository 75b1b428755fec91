"""Generates tests/golden/*.npz: seeded inputs with the oracle's outputs (fp32 restatement and
fp64 arbiter), written only after the two independent restatements (C literal loop, numpy closed
form) agree.  The reference itself cannot run here (Rust, no toolchain) so these pin the ORACLE,
and through it the CUDA path, against regressions; they are not reference outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import np_oracle as npo  # noqa: E402
import oracle_lib as ol  # noqa: E402

CASES = {  # name: (n_in, factor, pitch, W, seed, amplitude)
    "f2_p2_w256": (5000, 2.0, 2, 256, 11, 1.0),
    "f20_pm3_w1024": (2500, 20.0, -3, 1024, 12, 1.0),
    "f1p5_p3_w128": (3000, 1.5, 3, 128, 13, 0.5),
    "f8_p2_w2048": (5000, 8.0, 2, 2048, 14, 1.0),
    "f1_p1_w16384_short": (20000, 1.0, 1, 16384, 15, 1.0),
    "f3_pm2_w512": (3001, 3.0, -2, 512, 16, 1.0),
}


def make_input(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    x = 0.25 * (rng.random((2, n)) * 2 - 1) + 0.25 * np.sin(2 * np.pi * (200 + 3000 * t) * t)[None, :]
    return x.astype(np.float32)


def main():
    for name, (n, f, p, W, seed, amp) in CASES.items():
        x = make_input(n, seed)
        o64 = ol.process_all(x, f, p, W, seed=seed, amplitude=amp, precision=64)
        byp = ol.process_all(x, f, p, W, amplitude=amp, mode=ol.RO_PHASE_BYPASS)
        for c in range(2):
            r64 = npo.stretch_channel(x[c], f, p, W, amplitude=amp, seed=seed, channel=c, dtype=np.float64)
            assert np.abs(r64 - o64[c]).max() < 5e-7, name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, out_f64=o64, bypass=byp,
                            params=np.array([n, f, p, W, seed, amp], np.float64))
        print(name, o64.shape, "ok")


if __name__ == "__main__":
    main()
