"""Parity of the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs.  Bit-exact for all integer / index work (bypass mode); PCM within the north star's
tolerance (1e-4 max-abs, SNR > 100 dB) against the oracle's fp64 arbiter otherwise."""
import numpy as np
import pytest

import oracle_lib as ol
import rocoder_b200 as rb
from cases import LARGE_CASES, SMALL_CASES

pytestmark = pytest.mark.gpu

TOL_MAX_ABS = 1.0e-4   # test_utils.rs:4 / BASELINE.json north_star
TOL_SNR_DB = 100.0


def _input(n, channels=2, seed=1):
    rng = np.random.default_rng(seed)
    return ((rng.random((channels, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


def _gpu(x, f, p, W, seed=0, flags=0, amplitude=1.0):
    with rb.Engine(x.shape[0], 48000, W, f, amplitude, p, seed=seed, flags=flags) as e:
        return e.process_all(x)


@pytest.mark.parametrize("n,f,p,W", SMALL_CASES + LARGE_CASES)
def test_bypass_bit_exact(n, f, p, W):
    x = _input(n)
    want = ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS)
    got = _gpu(x, f, p, W, flags=rb.RCG_FLAG_BYPASS_RESYNTH)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,f,p,W", SMALL_CASES + LARGE_CASES)
def test_pcm_within_tolerance_of_oracle(n, f, p, W):
    x = _input(n)
    ref = ol.process_all(x, f, p, W, seed=1234, precision=64)
    got = _gpu(x, f, p, W, seed=1234)
    assert got.shape == ref.shape and np.isfinite(got).all()
    assert np.abs(got - ref).max() <= TOL_MAX_ABS
    if ref.any():
        assert ol.snr_db(ref, got) > TOL_SNR_DB


def test_gpu_equals_host_emulation_of_the_same_kernel():
    """same code, same tables: only libm (sincos) and FMA contraction differ"""
    n, f, p, W = 20000, 4.0, 2, 1024
    x = _input(n)
    a = _gpu(x, f, p, W, seed=77)
    b = ol.emul_process_all(x, f, p, W, seed=77)
    assert np.abs(a - b).max() < 2e-6


def test_amplitude_and_eight_channels():
    x = _input(9000, channels=8, seed=4)
    ref = ol.process_all(x, 2.0, 1, 1024, seed=9, amplitude=0.37, precision=64)
    got = _gpu(x, 2.0, 1, 1024, seed=9, amplitude=0.37)
    assert np.abs(got - ref).max() <= TOL_MAX_ABS and ol.snr_db(ref, got) > TOL_SNR_DB
    assert not np.array_equal(got[0], got[1])


def test_homogeneity_exact():
    """every stage is positively homogeneous and scaling by 2 is exact in binary fp:
    stretch(2x) == 2 stretch(x) bit for bit (a size-independent property)"""
    x = _input(50000, seed=6)
    a = _gpu(x, 3.0, 2, 4096, seed=5)
    b = _gpu(x * np.float32(2.0), 3.0, 2, 4096, seed=5)
    assert np.array_equal(b, a * np.float32(2.0))


def test_zero_input_gives_zero_output():
    out = _gpu(np.zeros((2, 40000), np.float32), 2.0, -2, 2048, seed=1)
    assert not out.any()


def test_deterministic_and_seed_sensitive():
    x = _input(30000)
    a = _gpu(x, 2.0, 1, 2048, seed=1)
    b = _gpu(x, 2.0, 1, 2048, seed=1)
    c = _gpu(x, 2.0, 1, 2048, seed=2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_random_seed_flag_matches_reference_statistics():
    """thread_rng mode (fft.rs:64) can only be checked statistically: zero mean, and the same
    output power as a seeded run (phases uniform on [0, PI), F2)"""
    x = _input(200000, seed=8)
    a = _gpu(x, 2.0, 1, 1024, seed=1)
    b = _gpu(x, 2.0, 1, 1024, flags=rb.RCG_FLAG_RANDOM_SEED)
    assert not np.array_equal(a, b)
    assert abs(b.mean()) < 1e-3
    assert abs((b ** 2).mean() / (a ** 2).mean() - 1.0) < 0.02


@pytest.mark.parametrize("f,p,W", [(4.0, 2, 512), (3.0, -3, 1024), (2.0, 1, 16384)])
def test_segments_on_device(f, p, W):
    """rcg_process_segment: halo recompute, tail chaining and deferred boundary + rcg_finish_boundary
    all reproduce the one-shot result bit for bit"""
    import torch
    n = 40 * W
    x = _input(n, seed=9)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=5) as e:
        whole = e.process_all(x)
        calls, _, out_len = e.expected_counts(n)
        pl = e.plan
        k, step, olen, H = pl["frames_per_call"], pl["sample_step_len"], pl["out_len_per_call"], W // 2
        cuts = [0, calls // 3, 2 * calls // 3 + 1, calls]
        dx = torch.from_numpy(x).cuda()
        stream = torch.cuda.current_stream().cuda_stream
        # (a) halo
        out = torch.full((2, out_len), float("nan"), device="cuda")
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            s0 = max(c0 * k - 1, 0) * step
            s1 = min(n, (c1 * k - 1) * step + W)
            seg = dx[:, s0:s1].contiguous()
            e.process_segment(seg, seg.stride(0), s0, seg.shape[1], out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                              stream=stream)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), whole)
        # (b) tail chain
        out = torch.full((2, out_len), float("nan"), device="cuda")
        tails = [torch.zeros(2, H, device="cuda") for _ in cuts]
        for i, (c0, c1) in enumerate(zip(cuts[:-1], cuts[1:])):
            e.process_segment(dx, dx.stride(0), 0, n, out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                              tail_in=tails[i] if c0 else None, tail_out=tails[i + 1], stream=stream)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), whole)
        # (c) deferred boundary
        out = torch.full((2, out_len), float("nan"), device="cuda")
        heads = [torch.zeros(2, H, device="cuda") for _ in cuts]
        tails = [torch.zeros(2, H, device="cuda") for _ in cuts]
        for i, (c0, c1) in enumerate(zip(cuts[:-1], cuts[1:])):
            e.process_segment(dx, dx.stride(0), 0, n, out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                              tail_out=tails[i], head_out=heads[i], defer_first=c0 > 0, stream=stream)
        for i, c0 in enumerate(cuts[1:-1], start=1):
            e.finish_boundary(out[:, c0 * olen:], out.stride(0), c0, heads[i], tails[i - 1], stream=stream)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), whole)


@pytest.mark.parametrize("f,p,W,n", [(2.0, 1, 1024, 30000), (8.0, 2, 512, 9000), (20.0, -3, 2048, 50000)])
def test_streaming_next_window_equals_batch(f, p, W, n):
    """Stretcher::next_window one window at a time (ragged chunks, then sender dropped) returns
    exactly the batched result and stops where the reference's processor stops"""
    x = _input(n, channels=1, seed=3)
    whole = _gpu(x, f, p, W, seed=8)
    rng = np.random.default_rng(0)
    with rb.Engine(1, 48000, W, f, 1.0, p, seed=8, max_lookahead_calls=3) as e:
        wins, pos = [], 0
        while pos < n:
            m = int(rng.integers(1, 3 * W))
            e.send(0, x[0, pos:pos + m])
            pos += m
            while True:
                try:
                    wins.append(e.next_window(0))
                except rb.WouldBlock:
                    break
            assert not e.is_done(0)
        e.close_input(0)
        while not e.is_done(0):
            wins.append(e.next_window(0))
    got = np.concatenate(wins)
    assert np.array_equal(got, whole[0])


def test_stretcher_processor_mirror():
    """the reference's wiring (main.rs:133-160): one Stretcher per channel fed one chunk, a
    StretcherProcessor thread, AudioBus::into_audio -- equals the batched engine"""
    n, f, p, W = 30000, 2.0, 2, 1024
    x = _input(n, seed=11)
    spec = rb.AudioSpec(2, 44100)
    stretchers = []
    for c in range(2):
        tx, rx = rb.unbounded()
        st = rb.Stretcher(spec, rx, f, 1.0, p, W, 1.0, None, channel=c, seed=21)
        tx.send(x[c])
        tx.drop()
        stretchers.append(st)
    proc, bus = rb.StretcherProcessor.new(stretchers, int(n * f))
    proc.start()
    audio = bus.into_audio()
    proc.join()
    whole = _gpu(x, f, p, W, seed=21)
    for c in range(2):
        assert np.array_equal(audio[c], whole[c])


def test_stretcher_ensure_input_samples_available():
    """stretcher.rs:145-161 against the host mirror"""
    spec = rb.AudioSpec(2, 44100)
    tx, rx = rb.unbounded()
    s = rb.Stretcher(spec, rx, 1.0, 1.0, 1, 1024, 1.0, None)
    tx.drop()
    s.ensure_input_samples_available(4)
    assert s.done is True and s.input_buf_len() == 4
    tx, rx = rb.unbounded()
    s = rb.Stretcher(spec, rx, 1.0, 1.0, 1, 1024, 1.0, None)
    tx.send(np.array([1.0, 2.0, 3.0], np.float32))
    tx.send(np.array([4.0, 5.0], np.float32))
    s.ensure_input_samples_available(4)
    assert s.done is False and s.input_buf_len() == 5


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (10 min stereo 48 kHz, -f 8 -p 2 -w 16384): output length,
    finiteness, chunk invariance on a slice, and the power the random-phase fold predicts"""
    from rocoder_b200.signals import noise_chirp
    n = 28_800_000
    x = noise_chirp(n, 48000, 2, seed=0)
    with rb.Engine(2, 48000, 16384, 8.0, 1.0, 2, seed=1234) as e:
        out = e.process_all(x)
        assert out.shape == (2, 230_277_120)
        assert np.isfinite(out[:, ::97]).all()
        # a middle slice recomputed as its own segment is bit-identical
        import torch
        c0, nc = 7000, 8
        s0 = (c0 * 4 - 1) * 512
        s1 = ((c0 + nc) * 4 - 1) * 512 + 16384
        seg = torch.from_numpy(x[:, s0:s1].copy()).cuda()
        o = torch.zeros(2, nc * 16384, device="cuda")
        e.process_segment(seg, seg.stride(0), s0, seg.shape[1], o, o.stride(0), c0, nc,
                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy(), out[:, c0 * 16384:(c0 + nc) * 16384])
    # oracle on the first calls of channel 0
    head = ol.process_all(x[:1, :16384 * 3], 8.0, 2, 16384, seed=1234, precision=64)
    m = 16384 * 8  # calls whose frames never touch the oracle's zero padding
    assert np.abs(out[0, :m] - head[0, :m]).max() <= TOL_MAX_ABS


def test_odd_output_stride_with_negative_pitch():
    """p <= -2 emits (S-1)*|p| samples per call (odd for W=4096, p=-3): segments then start at odd,
    4-byte-aligned offsets; only the direct-store modes (pitch 1 / 2) need 8-byte alignment"""
    import torch
    n, f, p, W = 60000, 3.0, -3, 4096
    x = _input(n, seed=2)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=5) as e:
        whole = e.process_all(x)
        calls, _, out_len = e.expected_counts(n)
        olen = e.plan["out_len_per_call"]
        assert olen % 2 == 1
        dx = torch.from_numpy(x).cuda()
        out = torch.full((2, out_len), float("nan"), device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        e.process_segment(dx, dx.stride(0), 0, n, out, out.stride(0), 0, 1, stream=st)
        e.process_segment(dx, dx.stride(0), 0, n, out[:, olen:], out.stride(0), 1, calls - 1, stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), whole)
    with rb.Engine(2, 48000, W, 2.0, 1.0, 1, seed=5) as e:   # pitch 1 still insists on alignment
        out = torch.zeros((2, 8 * W + 1), device="cuda")
        with pytest.raises(rb.RcgError):
            e.process_segment(dx, dx.stride(0), 0, n, out[:, 1:], out.stride(0), 0, 2, stream=st)


def test_window_65536_streaming_and_segments():
    """W = 65536 (2-CTA cluster per frame): streaming next_window and deferred-boundary segments equal
    the batched run bit for bit"""
    import torch
    n, f, p, W = 600000, 2.0, 2, 65536
    x = _input(n, seed=12)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=5) as e:
        whole = e.process_all(x)
        calls, _, out_len = e.expected_counts(n)
        olen, H = e.plan["out_len_per_call"], W // 2
        dx = torch.from_numpy(x).cuda()
        st = torch.cuda.current_stream().cuda_stream
        out = torch.full((2, out_len), float("nan"), device="cuda")
        cuts = [0, calls // 2, calls]
        heads = [torch.zeros(2, H, device="cuda") for _ in cuts]
        tails = [torch.zeros(2, H, device="cuda") for _ in cuts]
        for i, (c0, c1) in enumerate(zip(cuts[:-1], cuts[1:])):
            e.process_segment(dx, dx.stride(0), 0, n, out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                              tail_out=tails[i], head_out=heads[i], defer_first=c0 > 0, stream=st)
        e.finish_boundary(out[:, cuts[1] * olen:], out.stride(0), cuts[1], heads[1], tails[0], stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), whole)
    with rb.Engine(1, 48000, W, f, 1.0, p, seed=5, max_lookahead_calls=2) as e:
        e.send(0, x[0])
        e.close_input(0)
        wins = []
        while not e.is_done(0):
            wins.append(e.next_window(0))
    assert np.array_equal(np.concatenate(wins), whole[0])


def test_custom_window_vector():
    """Stretcher::new takes `window: Vec<f32>` (stretcher.rs:36); the reference's own tests pass a
    rectangular one (stretcher.rs:174, windows.rs:12-14)"""
    n, f, p, W = 9000, 2.0, 1, 1024
    x = _input(n, channels=1, seed=4)
    rect = np.ones(W, np.float32)
    s = ol.OracleStretcher(f, p, W, window=rect, precision=64)
    s.set_philox(3, 0)
    s.send(x[0]); s.close()
    ref = []
    while not s.is_done():
        ref.append(s.next_window())
    ref = np.concatenate(ref)
    with rb.Engine(1, 44100, W, f, 1.0, p, seed=3) as e:
        e.set_window(rect)
        got = e.process_all(x)[0]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-3 and ol.snr_db(ref, got) > 100.0   # rectangular frames are ~10x louder
    tx, rx = rb.unbounded()
    st = rb.Stretcher(rb.AudioSpec(1, 44100), rx, f, 1.0, p, rect, 1.0, None, seed=3)
    tx.send(x[0]); tx.drop()
    wins = []
    while not st.is_done():
        wins.append(st.next_window())
    assert np.array_equal(np.concatenate(wins), got)


@pytest.mark.gpu
@pytest.mark.parametrize("pitch", [1, -3, 2])
def test_custom_window_at_w16384_reaches_the_tensor_memory_image(pitch):
    """W = 16384: the full-rate launches read window and envelope from the Tensor Memory image (rcg_common.h:
    window_in_tmem), so rcg_set_window has to rebuild that image -- an asymmetric window tells a stale one apart.
    pitch 2 (half-rate path, global tables) runs beside it."""
    n, f, W = 120000, 2.0, 16384
    x = _input(n, channels=2, seed=9)
    win = (0.25 + 0.75 * np.linspace(0.0, 1.0, W, dtype=np.float64) ** 2).astype(np.float32)   # not symmetric
    got_default = None
    with rb.Engine(2, 48000, W, f, 1.0, pitch, seed=5) as e:
        got_default = e.process_all(x)
    with rb.Engine(2, 48000, W, f, 1.0, pitch, seed=5) as e:
        e.set_window(win)
        got = e.process_all(x)
    assert not np.array_equal(got, got_default)
    for c in range(2):
        s = ol.OracleStretcher(f, pitch, W, window=win, precision=64)
        s.set_philox(5, c)
        s.send(x[c]); s.close()
        ref = []
        while not s.is_done():
            ref.append(s.next_window())
        ref = np.concatenate(ref)
        assert got[c].shape == ref.shape
        assert np.abs(got[c] - ref).max() <= 1e-4 * max(1.0, float(np.abs(ref).max())) and ol.snr_db(ref, got[c]) > 100.0
