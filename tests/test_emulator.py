"""The kernel's own code (rocoder_b200/csrc/rcg_frame.h) executed thread-by-thread on the host
against the oracle: catches indexing / pairing / phase-slot / OLA / resample bugs before any GPU
time is spent.  The GPU tests repeat these checks through the C ABI on the device."""
import numpy as np
import pytest

import oracle_lib as ol
from cases import LARGE_CASES, SMALL_CASES


def _input(n, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


@pytest.mark.parametrize("n,f,p,W", SMALL_CASES + LARGE_CASES[:3] + LARGE_CASES[-3:])
def test_bypass_bit_exact(n, f, p, W):
    """all integer indexing (frame cursor, OLA blocks, decimation / lerp positions) is bit-exact"""
    x = _input(n)
    want = ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS)
    got = ol.emul_process_all(x, f, p, W, bypass=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,f,p,W", SMALL_CASES + LARGE_CASES[:2])
def test_full_path_within_tolerance(n, f, p, W):
    """north star: PCM within 1e-4 max-abs or SNR > 100 dB of the reference's fp32 (here: of the
    oracle's fp64 arbiter, which is the stricter comparison)"""
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=1234, precision=64)
    got = ol.emul_process_all(x, f, p, W, seed=1234, run_len=7)
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-4
    if ref.any():
        assert ol.snr_db(ref, got) > 100.0


def test_run_length_and_halo_do_not_change_results():
    n, f, p, W = 20000, 4.0, 2, 512
    x = _input(n)
    a = ol.emul_process_all(x, f, p, W, seed=3, run_len=1)
    b = ol.emul_process_all(x, f, p, W, seed=3, run_len=11)
    c = ol.emul_process_all(x, f, p, W, seed=3, run_len=10 ** 6)
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("f,p,W", [(4.0, 2, 512), (3.0, -3, 1024), (2.0, 1, 256)])
def test_segments_halo_tail_and_deferred_boundary(f, p, W):
    """time sharding: three segments computed independently equal the one-shot result bit for bit,
    whether the carried half window is recomputed (halo), passed in, or summed in afterwards"""
    n = 30000
    x = _input(n, seed=9)
    calls, _, out_len = ol.expected_counts(n, f, p, W)
    whole = ol.emul_process_all(x, f, p, W, seed=5)
    cuts = [0, calls // 3, 2 * calls // 3 + 1, calls]
    c = ol.constants(f, p, W)
    k, step, olen, H = c["frames_per_call"], c["sample_step_len"], c["out_len_per_call"], W // 2
    # (a) halo recompute: each segment only sees its own slice of the input (+ halo frame)
    out = np.full_like(whole, np.nan)
    for c0, c1 in zip(cuts[:-1], cuts[1:]):
        s0 = max(c0 * k - 1, 0) * step
        s1 = min(n, (c1 * k - 1) * step + W)
        seg_in = np.ascontiguousarray(x[:, s0:s1])
        ol.emul_segment(seg_in, s0, out[:, c0 * olen:], f, p, W, c0, c1 - c0, seed=5)
    assert np.array_equal(out, whole)
    # (b) tail passed from segment to segment
    out = np.full_like(whole, np.nan)
    tail = None
    for c0, c1 in zip(cuts[:-1], cuts[1:]):
        s0 = c0 * k * step
        seg_in = np.ascontiguousarray(x[:, s0:])
        tail_out = np.zeros((2, H), np.float32)
        ol.emul_segment(seg_in, s0, out[:, c0 * olen:], f, p, W, c0, c1 - c0, seed=5, tail_in=tail, tail_out=tail_out)
        tail = tail_out
    assert np.array_equal(out, whole)
    # (c) deferred first block + boundary fix-up (the multi-GPU exchange)
    out = np.full_like(whole, np.nan)
    heads, tails = [], []
    for c0, c1 in zip(cuts[:-1], cuts[1:]):
        s0 = c0 * k * step
        seg_in = np.ascontiguousarray(x[:, s0:])
        head, tail_out = np.zeros((2, H), np.float32), np.zeros((2, H), np.float32)
        ol.emul_segment(seg_in, s0, out[:, c0 * olen:], f, p, W, c0, c1 - c0, seed=5, tail_out=tail_out,
                        head_out=head, defer_first=c0 > 0)
        heads.append(head)
        tails.append(tail_out)
    for i, c0 in enumerate(cuts[1:-1], start=1):
        ol.emul_boundary(out[:, c0 * olen:], f, p, W, c0, heads[i], tails[i - 1])
    assert np.array_equal(out, whole)


def test_channel_index_keys_the_phases():
    x = _input(6000)
    x[1] = x[0]
    out = ol.emul_process_all(x, 2.0, 1, 256, seed=1)
    assert not np.array_equal(out[0], out[1])
    solo = np.full((1, out.shape[1]), np.nan, np.float32)
    calls = ol.expected_counts(6000, 2.0, 1, 256)[0]
    ol.emul_segment(np.ascontiguousarray(x[1:2]), 0, solo, 2.0, 1, 256, 0, calls, seed=1, first_channel=1)
    assert np.array_equal(solo[0], out[1])


@pytest.mark.parametrize("p", [2, 4, 6])
@pytest.mark.parametrize("W", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_half_rate_synthesis_even_pitch(p, W):
    """even pitch multiples run the synthesis side at half size (only y[2n] survives the
    decimation, resampler.rs:15-18): same results as the full-rate path up to rounding, for every
    FFT plan shape, bit-exact indexing, and within tolerance of the oracle"""
    n = 12 * W + 37
    f = 2.0
    x = _input(n, seed=W + p)
    byp = ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS)
    assert np.array_equal(ol.emul_process_all(x, f, p, W, bypass=True), byp)
    hr = ol.emul_process_all(x, f, p, W, seed=5)
    ol.emulator().rcg_emul_disable_half_rate(1)
    try:
        full = ol.emul_process_all(x, f, p, W, seed=5)
    finally:
        ol.emulator().rcg_emul_disable_half_rate(0)
    assert np.abs(hr - full).max() < 2e-6
    ref = ol.process_all(x, f, p, W, seed=5, precision=64)
    assert np.abs(hr - ref).max() <= 1e-4 and ol.snr_db(ref, hr) > 100.0


@pytest.mark.parametrize("n,f,p,W", [(30000, 4.0, 1, 2048), (40000, 8.0, 2, 4096), (70000, 8.0, 2, 16384),
                                     (9000, 20.0, -3, 1024), (6000, 2.5, 5, 512), (9000, 2.0, 4, 1024),
                                     (90000, 3.0, 1, 32768), (3000, 2.0, 2, 64)])
def test_phases_are_race_free(n, f, p, W):
    """no thread reads, within one phase, what another thread wrote in that phase: executing the
    emulated threads in descending instead of ascending order changes nothing (full path, bypass path)"""
    x = _input(n, seed=W + 1)
    E = ol.emulator()
    a = ol.emul_process_all(x, f, p, W, seed=5)
    b0 = ol.emul_process_all(x, f, p, W, bypass=True)
    E.rcg_emul_reverse_threads(1)
    try:
        b = ol.emul_process_all(x, f, p, W, seed=5)
        b1 = ol.emul_process_all(x, f, p, W, bypass=True)
    finally:
        E.rcg_emul_reverse_threads(0)
    assert np.array_equal(a, b) and np.array_equal(b0, b1)


@pytest.mark.parametrize("f,p,n", [(2.0, 1, 300000), (4.0, 2, 200000), (3.0, -3, 200000), (1.0, 3, 150000)])
def test_window_65536_clusters(f, p, n):
    """W = 65536 splits a frame across a 2-CTA thread-block cluster (even / odd bins, DSMEM exchange of the outer
    radix-2 stage): bit-exact indexing, PCM within tolerance, order-independent phases"""
    W = 65536
    x = _input(n, seed=p + 7)
    assert np.array_equal(ol.emul_process_all(x, f, p, W, bypass=True), ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS))
    ref = ol.process_all(x, f, p, W, seed=5, precision=64)
    got = ol.emul_process_all(x, f, p, W, seed=5, run_len=2)
    assert np.abs(got - ref).max() <= 1e-4 and ol.snr_db(ref, got) > 100.0
    ol.emulator().rcg_emul_reverse_threads(1)
    try:
        again = ol.emul_process_all(x, f, p, W, seed=5, run_len=3)
    finally:
        ol.emulator().rcg_emul_reverse_threads(0)
    assert np.array_equal(got, again)


def test_window_65536_segments():
    f, p, W, n = 2.0, 1, 65536, 500000
    x = _input(n, seed=3)
    calls, _, out_len = ol.expected_counts(n, f, p, W)
    whole = ol.emul_process_all(x, f, p, W, seed=5)
    c = ol.constants(f, p, W)
    k, step, olen, H = c["frames_per_call"], c["sample_step_len"], c["out_len_per_call"], W // 2
    cuts = [0, calls // 2, calls]
    out = np.full_like(whole, np.nan)
    heads, tails = [], []
    for c0, c1 in zip(cuts[:-1], cuts[1:]):
        s0 = c0 * k * step
        head, tail_out = np.zeros((2, H), np.float32), np.zeros((2, H), np.float32)
        ol.emul_segment(np.ascontiguousarray(x[:, s0:]), s0, out[:, c0 * olen:], f, p, W, c0, c1 - c0, seed=5,
                        tail_out=tail_out, head_out=head, defer_first=c0 > 0)
        heads.append(head)
        tails.append(tail_out)
    ol.emul_boundary(out[:, cuts[1] * olen:], f, p, W, cuts[1], heads[1], tails[0])
    assert np.array_equal(out, whole)
