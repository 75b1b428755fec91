"""Trust in the oracle beyond the reference's golden vectors: two independent
restatements (literal C deque loop + own Stockham FFT vs numpy closed form + pocketfft), an fp64
arbiter, Philox known answers, the closed-form call counts, and the statistics the reference's
quirks imply."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
from cases import SMALL_CASES

sys.path.insert(0, os.path.join(ol.ROOT, "oracle"))
import np_oracle as npo  # noqa: E402


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10"""
    L = ol.oracle()
    u4, u2 = C.c_uint32 * 4, C.c_uint32 * 2
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        out = u4()
        L.ro_philox4x32_10(u4(*ctr), u2(*key), out)
        assert list(out) == want
        got = npo.philox4x32_10(*[np.array([c], np.uint32) for c in ctr], key[0], key[1])
        assert [int(g[0]) for g in got] == want


def test_phase_mapping_range_and_slots():
    """fft.rs:13: phases are uniform on [0, PI), never PI; every bin has its own Philox word."""
    L = ol.oracle()
    assert L.ro_phase_from_u32(0) == 0.0
    top = L.ro_phase_from_u32(0xFFFFFFFF)
    assert top < np.float32(np.pi) and top > 3.14159
    for W in (8, 64, 1024):
        a = npo.phases_for_frame(0x1234567890ABCDEF, 3, (1 << 33) + 5, W)
        b = np.array([L.ro_phase(0x1234567890ABCDEF, 3, (1 << 33) + 5, k, W) for k in range(W)], np.float32)
        assert np.array_equal(a, b)
        buf = np.zeros(W, np.float32)
        L.ro_phase_fill(C.c_uint64(0x1234567890ABCDEF), 3, C.c_uint64((1 << 33) + 5), W, buf.ctypes.data_as(ol.PF))
        assert np.array_equal(buf, b)
        assert len(set(b.tolist())) > W * 0.9  # independent draws, no accidental sharing of words


def test_product_phase_matches_oracle():
    import rocoder_b200 as rb
    L = ol.oracle()
    for W in (64, 16384):
        for b in (0, 1, W // 4, W // 4 + 1, W // 2 - 1, W // 2, W // 2 + 1, 3 * W // 4, W - 1):
            assert rb.phase(99, 1, 12345, b, W) == L.ro_phase(99, 1, 12345, b, W)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 512, 4096, 16384, 12, 100])
def test_fft_against_numpy(n):
    L = ol.oracle()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for inverse in (0, 1):
        a = np.empty(2 * n)
        a[0::2], a[1::2] = x.real, x.imag
        L.ro_fft_f64(a.ctypes.data_as(C.c_void_p), C.c_size_t(n), inverse)
        want = np.fft.ifft(x) * n if inverse else np.fft.fft(x)
        assert np.abs((a[0::2] + 1j * a[1::2]) - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
        a32 = np.empty(2 * n, np.float32)
        a32[0::2], a32[1::2] = x.real, x.imag
        L.ro_fft_f32(a32.ctypes.data_as(C.c_void_p), C.c_size_t(n), inverse)
        assert np.abs((a32[0::2] + 1j * a32[1::2]) - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("n,f,p,W", SMALL_CASES)
def test_c_oracle_matches_numpy_restatement(n, f, p, W):
    rng = np.random.default_rng(1)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    assert ol.expected_counts(n, f, p, W) == npo.counts(n, f, p, W)
    o32 = ol.process_all(x, f, p, W, seed=1234)
    o64 = ol.process_all(x, f, p, W, seed=1234, precision=64)
    for c in range(2):
        r64 = npo.stretch_channel(x[c], f, p, W, seed=1234, channel=c, dtype=np.float64)
        assert r64.shape == o64[c].shape
        # same algorithm, different FFT + different libm window: agreement to ~1e-7 of a ~0.7 peak
        assert np.abs(r64 - o64[c]).max() < 5e-7
        assert np.abs(o32[c] - o64[c]).max() < 5e-6


def test_constants_table_from_survey():
    """the five BASELINE configs, verified there by literal emulation."""
    rows = [  # f, p, W -> S, step, k
        (1.0, 1, 16384, 16384, 8192, 2), (8.0, 2, 16384, 32768, 512, 4), (4.0, 1, 16384, 16384, 2048, 2),
        (20.0, -3, 16384, 5462, 1228, 1), (2.0, 1, 16384, 16384, 4096, 2),
    ]
    for f, p, W, S, step, k in rows:
        c = ol.constants(f, p, W)
        assert (c["samples_needed_per_window"], c["sample_step_len"], c["frames_per_call"]) == (S, step, k)
        assert c["corrected_amp_factor"] == 4.0
    assert ol.constants(20.0, -3, 16384)["out_len_per_call"] == 16383   # F5(b)
    assert ol.constants(8.0, -2, 16384)["out_len_per_call"] == 16382


def test_counts_table_from_survey():
    """BASELINE.md section 5 (per channel)"""
    rows = [(441000, 1.0, 1, 16384, 54, 442368), (28800000, 8.0, 2, 16384, 56220, 230277120),
            (28800000, 4.0, 1, 1024, 224994, 115196928), (28800000, 4.0, 1, 4096, 56244, 115187712),
            (28800000, 4.0, 1, 16384, 14056, 115146752), (28800000, 4.0, 1, 65536, 3510, 115015680),
            (172800000, 20.0, -3, 16384, 140705, 2305170015), (5760000, 2.0, 1, 16384, 1404, 11501568)]
    for n, f, p, W, frames, out in rows:
        _, fr, o = ol.expected_counts(n, f, p, W)
        assert (fr, o) == (frames, out)


@pytest.mark.parametrize("n,f,p,W", [(5000, 1.0, 1, 256), (7001, 3.0, -2, 512), (9000, 20.0, -3, 1024),
                                     (6000, 2.5, 5, 512), (300, 7.0, 2, 64)])
def test_closed_form_counts_equal_literal_loop(n, f, p, W):
    s = ol.OracleStretcher(f, p, W)
    s.set_bypass()
    s.send(np.zeros(n, np.float32))
    s.close()
    calls = 0
    while not s.is_done():  # stretcher_processor.rs:64-69
        s.next_window()
        calls += 1
    assert calls == ol.expected_counts(n, f, p, W)[0]


def test_identity_phase_reduces_to_windowed_ola():
    """with analysis phases kept, resynth(s) = s*w*w and the pipeline is plain OLA."""
    W, n = 256, 4000
    x = np.random.default_rng(3).standard_normal((1, n)).astype(np.float32) * 0.2
    out = ol.process_all(x, 1.0, 1, W, mode=ol.RO_PHASE_BYPASS)
    w, env = ol.hanning(W), ol.crossfade(W // 2)
    H = W // 2
    calls, frames, _ = ol.expected_counts(n, 1.0, 1, W)
    xp = np.concatenate([x[0], np.zeros(frames * H + W, np.float32)])
    prev = np.zeros(H, np.float32)
    want = []
    for j in range(frames):
        y = (xp[j * H: j * H + W] * w) * w
        want.append(((y[:H] + prev) * env) * np.float32(4.0))
        prev = y[H:]
    assert np.array_equal(out[0], np.concatenate(want))


def test_random_phase_statistics():
    """F1/F2: i.i.d. phases on [0, PI): E|(e^{i a} + e^{-i b})/2|^2 = 0.5 - 2/pi^2 = 0.2974 (not 0.5),
    so a white input of power P leaves resynth with power 0.2974 * mean(w^2 * ...)."""
    W = 1024
    rng = np.random.default_rng(5)
    ph = np.stack([npo.phases_for_frame(7, 0, j, W) for j in range(64)])
    assert 0.0 <= ph.min() and ph.max() < np.pi
    assert abs(ph.mean() - np.pi / 2) < 0.02
    a, b = ph[:, 1:W // 2], ph[:, W - 1:W // 2:-1]
    fold = np.abs((np.exp(1j * a) + np.exp(-1j * b)) / 2) ** 2
    assert abs(fold.mean() - (0.5 - 2 / np.pi ** 2)) < 0.01
    # output is zero mean
    x = rng.standard_normal((1, 60000)).astype(np.float32) * 0.1
    out = ol.process_all(x, 2.0, 1, W, seed=7)
    assert abs(out.mean()) < 1e-3


def test_device_phase_formula_is_the_oracles_for_every_mantissa():
    """rcg_philox.h, device branch of phase_from_u32: x = 1 + v 2^-23 (v = u >> 9 as the mantissa of a float in [1, 2)),
    phase = fma(x, PI, -PI).  x PI - PI is exactly v 2^-23 PI (48 significant bits: exact in float64), so the single
    rounding of the fused multiply-add gives the value the oracle's product ((u >> 9) * 2^-23) * PI gives -- checked
    here for all 2^23 mantissas against the oracle's own function on a sample and against its formula on all."""
    L = ol.oracle()
    v = np.arange(1 << 23, dtype=np.uint32)
    pi32 = np.float64(np.float32(np.pi))
    x = (v | np.uint32(0x3F800000)).view(np.float32).astype(np.float64)          # [1, 2)
    fused = (x * pi32 - pi32).astype(np.float32)                                   # exact in float64, rounded once
    product = (v.astype(np.float32) * np.float32(2.0 ** -23 * float(np.float32(np.pi)))).astype(np.float32)
    assert np.array_equal(fused, product)
    for vi in (0, 1, 2, 12345, (1 << 22) + 1, (1 << 23) - 1):
        assert fused[vi] == np.float32(L.ro_phase_from_u32(int(vi) << 9))
        assert fused[vi] == np.float32(L.ro_phase_from_u32((int(vi) << 9) | 0x1FF))
