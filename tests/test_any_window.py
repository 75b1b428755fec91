"""Windows that are not a power of two (reference README.md:100: any length is legal, fft.rs:27-29 plans it).  The
engine runs them through Bluestein's chirp-z identity on its power-of-two shared-memory passes
(rocoder_b200/csrc/rcg_frame_bluestein.h); the oracle's plain O(n^2) DFT (oracle_core.inc dft_naive) is the checker.
Above 8192 samples (and for powers of two above 65536) the convolutions run through HBM (rcg_big.h); there the
numpy restatement of the oracle (pocketfft in fp64, any length) is the checker.
CPU tests execute the kernel's own phases on the host emulator; GPU tests go through the C ABI."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import np_oracle as npo  # noqa: E402
import rocoder_b200 as rb
from rocoder_b200 import _abi

CASES = [(9000, 2.0, 1, 1000), (7000, 3.0, 2, 600), (12000, 4.0, -3, 1500), (3000, 1.0, 1, 66), (20000, 2.0, -2, 3000),
         (5000, 1.5, 3, 250), (12000, 8.0, 2, 4410)]
TOL_MAX_ABS, TOL_SNR_DB = 1.0e-4, 100.0


def _input(n, channels=2, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((channels, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


def test_plan_accepts_even_lengths_and_rejects_odd_ones():
    p = rb.plan(2.0, 1, 1000)
    assert p["window_len"] == 1000 and p["half_window_len"] == 500 and p["sample_step_len"] == 250
    assert rb.expected_counts(9000, 2.0, 1, 1000) == ol.expected_counts(9000, 2.0, 1, 1000)
    for bad in (1001, 63, 62, 8195, 2 ** 20 + 2, 2 ** 21):
        with pytest.raises(rb.RcgError) as ei:
            rb.plan(2.0, 1, bad)
        assert ei.value.status == _abi.RCG_ERR_WINDOW
    for good in (64, 66, 1000, 8190, 8192, 8194, 10000, 16384, 44100, 65536, 100000, 131072, 2 ** 20):
        rb.plan(2.0, 1, good)


@pytest.mark.parametrize("n,f,p,W", CASES[:6])
def test_emulated_chirp_z_path_matches_oracle(n, f, p, W):
    """the kernel's phases executed on the host: index path bit-exact, PCM within tolerance of the fp64 oracle, and
    independent of the thread order inside a phase"""
    x = _input(n)
    assert np.array_equal(ol.emul_process_all(x, f, p, W, bypass=True), ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS))
    ref = ol.process_all(x, f, p, W, seed=5, precision=64)
    got = ol.emul_process_all(x, f, p, W, seed=5, run_len=3)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL_MAX_ABS and ol.snr_db(ref, got) > TOL_SNR_DB
    ol.emulator().rcg_emul_reverse_threads(1)
    try:
        again = ol.emul_process_all(x, f, p, W, seed=5, run_len=7)
    finally:
        ol.emulator().rcg_emul_reverse_threads(0)
    assert np.array_equal(got, again)


def test_emulated_chirp_z_segments_and_boundary():
    f, p, W, n = 2.0, 1, 1000, 30000
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


@pytest.mark.gpu
@pytest.mark.parametrize("n,f,p,W", CASES)
def test_gpu_chirp_z_path_matches_oracle(n, f, p, W):
    x = _input(n)
    with rb.Engine(2, 48000, W, f, 1.0, p, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
        assert np.array_equal(e.process_all(x), ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS))
    ref = ol.process_all(x, f, p, W, seed=1234, precision=64)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1234) as e:
        got = e.process_all(x)
        assert e.launch_count() >= 1
    assert got.shape == ref.shape and np.isfinite(got).all()
    assert np.abs(got - ref).max() <= TOL_MAX_ABS and ol.snr_db(ref, got) > TOL_SNR_DB


@pytest.mark.gpu
def test_gpu_chirp_z_streaming_segments_devices_and_kernels():
    """everything that rides on a window size works for the chirp-z path too: streaming next_window, time shards
    behind one engine (exchange and halo), device segments; frequency kernels are refused with a clear status"""
    n, f, p, W = 60000, 2.0, 2, 1000
    x = _input(n, seed=7)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=3) as e:
        whole = e.process_all(x)
        with pytest.raises(rb.RcgError) as ei:
            e.push_local_device_frequency_kernel("__device__ float2 apply(size_t a, unsigned b, unsigned c, float2 x) { return x; }")
        assert ei.value.status == _abi.RCG_ERR_STATE
        with pytest.raises(rb.RcgError):
            e.push_frequency_kernel(lambda ms, s: s)
    for flags in (rb.RCG_FLAG_SHARD_ALWAYS, rb.RCG_FLAG_SHARD_ALWAYS | rb.RCG_FLAG_SHARD_HALO):
        with rb.Engine(2, 48000, W, f, 1.0, p, seed=3, devices=[0, 0, 0], flags=flags) as g:
            assert np.array_equal(g.process_all(x), whole)
            assert g.last_partition()["split"] == "time"
    with rb.Engine(1, 48000, W, f, 1.0, p, seed=3, max_lookahead_calls=3) as e:
        e.send(0, x[0, :20000]); e.send(0, x[0, 20000:]); e.close_input(0)
        wins = []
        while not e.is_done(0):
            wins.append(e.next_window(0))
    assert np.array_equal(np.concatenate(wins), whole[0])


# ---- long windows: chirp-z convolutions through HBM (rcg_big.h) ------------------------------------------------------
# (samples, factor, pitch, window): just above the in-SM limit; the sample rate as a window; a power of two above
# 65536; negative and even pitch multiples; enough frames to cross a batch boundary (1024 frames at W = 8194)
BIG_CASES = [(60000, 2.0, 1, 10000), (90000, 3.0, 2, 12000), (400000, 4.0, -3, 44100), (150000, 2.0, -2, 20000),
             (600000, 1.5, 1, 131072), (320000, 8.0, 2, 8194)]


def _np_oracle_all(x, f, p, W, seed, dtype=np.float64):
    return np.stack([npo.stretch_channel(x[c], f, p, W, seed=seed, channel=c, dtype=dtype) for c in range(x.shape[0])])


@pytest.mark.gpu
@pytest.mark.parametrize("n,f,p,W", BIG_CASES)
def test_gpu_long_window_path_matches_oracle(n, f, p, W):
    x = _input(n)
    with rb.Engine(2, 48000, W, f, 1.0, p, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
        assert np.array_equal(e.process_all(x), ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS))
    ref = _np_oracle_all(x, f, p, W, 1234)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=1234) as e:
        got = e.process_all(x)
        assert e.launch_count() >= 1
    assert got.shape == ref.shape and np.isfinite(got).all()
    assert np.abs(got - ref).max() <= TOL_MAX_ABS and ol.snr_db(ref, got) > TOL_SNR_DB


@pytest.mark.gpu
def test_gpu_long_window_streaming_segments_and_devices():
    n, f, p, W = 200000, 2.0, 2, 10000
    x = _input(n, seed=7)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=3) as e:
        whole = e.process_all(x)
        with pytest.raises(rb.RcgError) as ei:
            e.push_local_device_frequency_kernel("__device__ float2 apply(size_t a, unsigned b, unsigned c, float2 x) { return x; }")
        assert ei.value.status == _abi.RCG_ERR_STATE
    for flags in (rb.RCG_FLAG_SHARD_ALWAYS, rb.RCG_FLAG_SHARD_ALWAYS | rb.RCG_FLAG_SHARD_HALO):
        with rb.Engine(2, 48000, W, f, 1.0, p, seed=3, devices=[0, 0, 0], flags=flags) as g:
            assert np.array_equal(g.process_all(x), whole)
            assert g.last_partition()["split"] == "time"
    with rb.Engine(1, 48000, W, f, 1.0, p, seed=3, max_lookahead_calls=3) as e:
        e.send(0, x[0, :70000]); e.send(0, x[0, 70000:]); e.close_input(0)
        wins = []
        while not e.is_done(0):
            wins.append(e.next_window(0))
    assert np.array_equal(np.concatenate(wins), whole[0])


def test_np_oracle_agrees_with_c_oracle_on_a_window_that_is_not_a_power_of_two():
    """the checker of the long-window tests (pocketfft, fp64) against the literal C restatement (plain DFT, fp64):
    same algorithm, another FFT and another libm behind the f32 window -- agreement to ~5e-8 of a ~0.7 peak"""
    x = _input(9000)
    for (f, p, W) in ((2.0, 1, 1000), (3.0, -2, 600), (1.5, 3, 250), (2.0, 2, 66)):   # the last two: W / 2 odd
        a = ol.process_all(x, f, p, W, seed=5, precision=64)
        b = _np_oracle_all(x, f, p, W, 5)
        assert a.shape == b.shape and np.abs(a - b).max() < 5e-7 and ol.snr_db(a, b) > 140.0
