"""Frequency kernels used by the tests, in the C-ABI form of the reference's `apply`
(include/rocoder_b200.h; reference fft.rs:93, README.md:108-129)."""
import numpy as np


def _views(pin, n_bins, pout):
    a = np.ctypeslib.as_array(pin, shape=(2 * n_bins,))
    b = np.ctypeslib.as_array(pout, shape=(2 * n_bins,))
    return a, b


def amplitude_x2(elapsed_ms, pin, n_bins, pout):
    """README.md:121-129: every bin times 2"""
    a, b = _views(pin, n_bins, pout)
    b[:] = a * np.float32(2.0)
    return 0


def lowpass_asymmetric(elapsed_ms, pin, n_bins, pout):
    """keeps bins [0, n/8) only: NOT Hermitian-symmetric, so bins k and W-k get different magnitudes"""
    a, b = _views(pin, n_bins, pout)
    b[:] = 0
    b[: 2 * (n_bins // 8)] = a[: 2 * (n_bins // 8)]
    return 0


def tilt(elapsed_ms, pin, n_bins, pout):
    """bin-dependent gain, different for every one of the W bins, plus a phase rotation (discarded, fft.rs:67)"""
    a, b = _views(pin, n_bins, pout)
    z = a.view(np.complex64)
    g = (0.25 + np.arange(n_bins, dtype=np.float32) / np.float32(n_bins)).astype(np.float32)
    b.view(np.complex64)[:] = z * g * np.complex64(0.6 + 0.8j)
    return 0


def panics(elapsed_ms, pin, n_bins, pout):
    return 1
