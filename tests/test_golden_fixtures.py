"""Committed fixtures (tests/golden, made by make_golden.py): the oracle and the host emulation of
the kernel on CPU; the CUDA path through the C ABI on GPU."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as ol

FILES = sorted(glob.glob(os.path.join(ol.ROOT, "tests", "golden", "*.npz")))


def _load(path):
    d = np.load(path)
    n, f, p, W, seed, amp = d["params"]
    return d, int(n), float(f), int(p), int(W), int(seed), float(amp)


def test_fixtures_exist():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_fixture(path):
    d, n, f, p, W, seed, amp = _load(path)
    assert np.array_equal(ol.process_all(d["x"], f, p, W, seed=seed, amplitude=amp, mode=ol.RO_PHASE_BYPASS), d["bypass"])
    o64 = ol.process_all(d["x"], f, p, W, seed=seed, amplitude=amp, precision=64)
    assert np.abs(o64 - d["out_f64"]).max() < 1e-9  # libm sin/cos may differ in the last bit across hosts


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_kernel_emulation_matches_fixture(path):
    d, n, f, p, W, seed, amp = _load(path)
    assert np.array_equal(ol.emul_process_all(d["x"], f, p, W, amplitude=amp, bypass=True), d["bypass"])
    got = ol.emul_process_all(d["x"], f, p, W, seed=seed, amplitude=amp)
    assert np.abs(got - d["out_f64"]).max() <= 1e-4 and ol.snr_db(d["out_f64"], got) > 100.0


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_path_matches_fixture(path):
    import rocoder_b200 as rb
    d, n, f, p, W, seed, amp = _load(path)
    with rb.Engine(2, 48000, W, f, amp, p, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
        assert np.array_equal(e.process_all(d["x"]), d["bypass"])
    with rb.Engine(2, 48000, W, f, amp, p, seed=seed) as e:
        got = e.process_all(d["x"])
    assert np.abs(got - d["out_f64"]).max() <= 1e-4 and ol.snr_db(d["out_f64"], got) > 100.0
