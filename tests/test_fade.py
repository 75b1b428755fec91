"""the `-x` fade of the reference's live mixer (mixer.rs:132-190, math.rs:33-40).
Keyframe arithmetic on CPU, the elementwise kernel on GPU (bit-exact against the oracle)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import rocoder_b200 as rb


@pytest.mark.parametrize("fade_s,rate,total", [(1.0, 48000, 480000), (0.25, 44100, 100000), (3.0, 96000, 96000 * 7 + 13),
                                               (1.0, 44100, int(441000 * 1.0)), (2.5, 48000, 48000 * 4)])
def test_fade_keyframes_match_oracle(fade_s, rate, total):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    ol.oracle().ro_fade_keyframes(fade_s, rate, total, C.byref(a), C.byref(b), C.byref(c))
    assert rb.fade_keyframes(fade_s, rate, total) == (a.value, b.value, c.value)


def test_fade_curve_samples_from_reference_tests():
    """audio.rs:290-313: gains 0.5, 0.70710677, 0.8660254 a quarter, half, three quarters into a fade in"""
    L = ol.oracle()
    got = [L.ro_fade_gain(p, 1.0, 48000, 480000) for p in (12000, 24000, 36000)]
    assert np.allclose(got, [0.5, 0.70710677, 0.8660254], atol=1e-4)
    assert L.ro_fade_gain(0, 1.0, 48000, 480000) == 0.0 and L.ro_fade_gain(240000, 1.0, 48000, 480000) == 1.0
    assert L.ro_fade_gain(480000, 1.0, 48000, 480000) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("fade_s,rate,n", [(1.0, 48000, 480000), (0.25, 44100, 100001)])
def test_fade_kernel_bit_exact(fade_s, rate, n):
    import torch
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, n)).astype(np.float32)
    d = torch.from_numpy(x.copy()).cuda()
    rb.fade_in_out(d, fade_s, rate, n, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    L = ol.oracle()
    pos = list(range(0, int(fade_s * rate) + 5)) + list(range(n - int(fade_s * rate) - 5, n)) + [n // 2]
    gain = np.array([L.ro_fade_gain(p, fade_s, rate, n) for p in pos], np.float32)
    got = d.cpu().numpy()
    for c in range(2):
        assert np.array_equal(got[c, pos], x[c, pos] * gain)
    mid = slice(int(fade_s * rate) + 1, n - int(fade_s * rate) - 1)
    assert np.array_equal(got[:, mid], x[:, mid])


@pytest.mark.gpu
def test_fp32_peak_microbenchmark():
    t = rb.measure_fp32_peak()
    assert 30.0 < t < 90.0   # B200: 148 SMs x 128 lanes x 2 x ~1.9 GHz = ~72 TFLOP/s nominal
