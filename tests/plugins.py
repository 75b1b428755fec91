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


def half_gain(elapsed_ms, pin, n_bins, pout):
    a, b = _views(pin, n_bins, pout)
    b[:] = a * np.float32(0.5)
    return 0


class FailsAfter:
    """a stateful host kernel: half_gain for the first `good` frames, then it "panics" on every call (fft.rs:101-106:
    the kernel is dropped and the frame retried with the one beneath it)"""

    def __init__(self, good):
        self.good, self.calls = good, 0

    def __call__(self, elapsed_ms, pin, n_bins, pout):
        self.calls += 1
        if self.calls > self.good:
            return 1
        return half_gain(elapsed_ms, pin, n_bins, pout)


class GainThenTilt:
    """what the reference computes for the stack [tilt, FailsAfter(good)]: half_gain for `good` frames, tilt after"""

    def __init__(self, good):
        self.good, self.calls = good, 0

    def __call__(self, elapsed_ms, pin, n_bins, pout):
        self.calls += 1
        return half_gain(elapsed_ms, pin, n_bins, pout) if self.calls <= self.good else tilt(elapsed_ms, pin, n_bins, pout)


# device forms of the same kernels (CUDA source for NVRTC): gather form, then the bin-local form
DEVICE_X2 = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    float2 v = spectrum[bin];
    return make_float2(v.x * 2.0f, v.y * 2.0f);        // README.md:121-129
}
"""
DEVICE_TILT = """
float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {   // -default-device
    float g = 0.25f + (float)bin / (float)n_bins;
    float2 v = spectrum[bin];
    return make_float2((v.x * 0.6f - v.y * 0.8f) * g, (v.x * 0.8f + v.y * 0.6f) * g);
}
"""
DEVICE_LOWPASS = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    return bin < n_bins / 8 ? spectrum[bin] : make_float2(0.0f, 0.0f);
}
"""
LOCAL_X2 = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {
    return make_float2(x.x * 2.0f, x.y * 2.0f);        // README.md:121-129
}
"""
LOCAL_TILT = """
float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 v) {   // -default-device
    float g = 0.25f + (float)bin / (float)n_bins;
    return make_float2((v.x * 0.6f - v.y * 0.8f) * g, (v.x * 0.8f + v.y * 0.6f) * g);
}
"""
LOCAL_LOWPASS = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {
    return bin < n_bins / 8 ? x : make_float2(0.0f, 0.0f);
}
"""
