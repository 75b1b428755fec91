"""Every BASELINE.json config at its stated size on the GPU, checked the cheap way: the whole job runs
device-resident; then slices of it -- the first calls, calls in the middle, the last calls before the zero padding,
and for config 4 calls whose output index lies past 2^31 -- are (a) recomputed as their own segment (halo frame
recomputed, different work split, different buffer offsets) and compared bit for bit, and (b) recomputed by the
oracle in fp64, started mid-stream with the right phase counters, and compared within the north star's tolerance.
Config 4 also runs the index path (bypass) at full size against the oracle bit for bit."""
import numpy as np
import pytest

import oracle_lib as ol
import plugins
import rocoder_b200 as rb
from rocoder_b200.signals import noise_chirp_span_torch

pytestmark = pytest.mark.gpu

TOL_MAX_ABS = 1.0e-4   # BASELINE.json north_star / test_utils.rs:4
TOL_SNR_DB = 100.0
SEED = 1234


def _slices(calls, k, n_in, step, W, extra=()):
    """first calls, middle, the last calls whose frames lie inside the input, plus `extra`"""
    last_full = max(0, (n_in - W) // step // k - 1)      # last call that reads no zero padding
    cand = [0, calls // 2, max(0, last_full - 3)] + list(extra)
    return sorted(set(c for c in cand if c + 4 <= calls))


def _run_and_check(channels, rate, seconds, f, p, W, extra_slices=(), n_check=4, engine_setup=None, oracle_plugin=None,
                   flags=0, oracle_channels=(0,)):
    import torch
    n = rate * seconds
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    x = noise_chirp_span_torch(0, n, rate, channels, seed=0, device=dev, total_seconds=float(seconds))
    with rb.Engine(channels, rate, W, f, 1.0, p, seed=SEED, flags=flags) as e:
        if engine_setup:
            engine_setup(e)
        calls, frames, out_len = e.expected_counts(n)
        pl = e.plan
        k, step, olen = pl["frames_per_call"], pl["sample_step_len"], pl["out_len_per_call"]
        out = torch.empty((channels, out_len), dtype=torch.float32, device=dev)
        e.process_segment(x, x.stride(0), 0, n, out, out.stride(0), 0, calls, stream=st)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out[:, ::1009]).all())
        # the zero-padded last call is emitted in full (stretcher.rs:129-131, stretcher_processor.rs:64-68)
        assert out.shape[1] == calls * olen
        checked = []
        for c0 in _slices(calls, k, n, step, W, extra_slices):
            nc = n_check
            s0 = max(c0 * k - 1, 0) * step
            s1 = min(n, ((c0 + nc) * k - 1) * step + W)
            # (a) the same calls as their own segment: bit-identical
            seg = x[:, s0:s1].contiguous()
            o = torch.full((channels, nc * olen + 2), float("nan"), device=dev)
            # pitch 1 / 2 store float2 straight from the overlap-add registers (8-byte aligned output, even stride); the
            # resampled modes take any 4-byte aligned offset: give them an odd one
            o_view = o[:, :nc * olen] if p in (1, 2) else o[:, 1:1 + nc * olen]
            e.process_segment(seg, seg.stride(0), s0, seg.shape[1], o_view, o.stride(0), c0, nc, stream=st)
            torch.cuda.synchronize()
            big = out[:, c0 * olen:(c0 + nc) * olen]
            assert torch.equal(o_view, big), f"call {c0}: segment recomputation differs"
            # (b) the oracle, fp64, mid-stream
            if s1 - s0 >= ((c0 + nc) * k - 1) * step + W - s0:     # no zero padding inside the slice
                lead0 = max(c0 - 1, 0) * k * step
                for ch in oracle_channels:
                    span = x[ch, lead0:s1].cpu().numpy()
                    if flags & rb.RCG_FLAG_BYPASS_RESYNTH:
                        ref = ol.process_calls(span, lead0, f, p, W, SEED, ch, c0, nc, precision=32, bypass=True)
                        assert np.array_equal(big[ch].cpu().numpy(), ref), f"call {c0} channel {ch}: index path differs"
                    elif oracle_plugin is None:
                        ref = ol.process_calls(span, lead0, f, p, W, SEED, ch, c0, nc, precision=64)
                        got = big[ch].cpu().numpy()
                        assert np.abs(got - ref).max() <= TOL_MAX_ABS, f"call {c0} channel {ch}"
                        assert ol.snr_db(ref, got) > TOL_SNR_DB
                checked.append(c0)
        assert len(checked) >= 2
        return out, x, (calls, k, step, olen)


def test_config1_real_size_fully_against_the_oracle():
    """config 1: 10 s stereo 44.1 kHz, -f 1 -p 1 -w 16384 -- every sample of the job against the fp64 oracle, through
    the host-buffer call a file-to-file run makes"""
    from rocoder_b200.signals import sine_sweep
    x = sine_sweep(441_000, 44100, 2)
    with rb.Engine(2, 44100, 16384, 1.0, 1.0, 1, seed=SEED) as e:
        got = e.process_all(x)
    ref = ol.process_all(x, 1.0, 1, 16384, seed=SEED, precision=64, sample_rate=44100)
    assert got.shape == ref.shape == (2, 442_368)
    assert np.abs(got - ref).max() <= TOL_MAX_ABS and ol.snr_db(ref, got) > TOL_SNR_DB
    byp = ol.process_all(x, 1.0, 1, 16384, mode=ol.RO_PHASE_BYPASS, sample_rate=44100)
    with rb.Engine(2, 44100, 16384, 1.0, 1.0, 1, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
        assert np.array_equal(e.process_all(x), byp)


@pytest.mark.parametrize("W,out_len", [(1024, 115_196_928), (4096, 115_187_712), (16384, 115_146_752), (65536, 115_015_680)])
def test_config3_window_sweep_full_size(W, out_len):
    """config 3: 10 min stereo 48 kHz, -f 4, every window of the sweep (SURVEY 8d output lengths)"""
    out, _, _ = _run_and_check(2, 48000, 600, 4.0, 1, W, oracle_channels=(0, 1) if W <= 4096 else (1,))
    assert out.shape == (2, out_len)


@pytest.mark.timeout(900)
def test_config4_one_hour_past_2_to_the_31():
    """config 4: 1 h stereo 48 kHz, -f 20 -p -3: 140 705 calls per channel, each dropping 2730 finished samples
    (stretcher.rs:91, 108-112) and emitting (S-1)*3 = 16383 (resampler.rs:20-35); output index reaches 2 305 170 015"""
    import torch
    out, x, (calls, k, step, olen) = _run_and_check(2, 48000, 3600, 20.0, -3, 16384, extra_slices=(131_080, 135_000, 140_000),
                                                    oracle_channels=(0, 1))
    assert (calls, k, step, olen) == (140_705, 1, 1228, 16383)
    assert out.shape == (2, 2_305_170_015) and 135_000 * olen > 2 ** 31
    del out
    torch.cuda.empty_cache()
    # the index path at full size, bit for bit (lerp positions, dropped samples, offsets past 2^31)
    _run_and_check(2, 48000, 3600, 20.0, -3, 16384, extra_slices=(131_080, 140_000), flags=rb.RCG_FLAG_BYPASS_RESYNTH)


@pytest.mark.parametrize("form", ["fused", "gather", "host"])
def test_config5_eight_channels_with_the_x2_kernel(form):
    """config 5: 7.1 (8 channels) 96 kHz, 60 s, -f 2 with the README's amplitude-scale kernel (README.md:121-129) in
    each of the three forms the engine takes it in; slices against the oracle running the same kernel"""
    import torch
    n, rate, f, p, W = 60 * 96000, 96000, 2.0, 1, 16384
    dev = torch.device("cuda", 0)
    x = noise_chirp_span_torch(0, n, rate, 8, seed=0, device=dev, total_seconds=60.0)
    xh = x.cpu().numpy()
    with rb.Engine(8, rate, W, f, 1.0, p, seed=SEED) as e:
        if form == "fused":
            e.push_local_device_frequency_kernel(plugins.LOCAL_X2)
        elif form == "gather":
            e.push_device_frequency_kernel(plugins.DEVICE_X2)
        else:
            e.push_frequency_kernel(lambda ms, spec: spec * np.float32(2.0))
        got = e.process_all(xh)
        calls, frames, out_len = e.expected_counts(n)
        k, step, olen = e.plan["frames_per_call"], e.plan["sample_step_len"], e.plan["out_len_per_call"]
    assert got.shape == (8, 11_501_568) and frames == 1404 and np.isfinite(got).all()
    for c0 in (0, calls // 2, calls - 8):
        lead0 = max(c0 - 1, 0) * k * step
        s1 = ((c0 + 3) * k - 1) * step + W
        for ch in (0, 7):
            ref = ol.process_calls(xh[ch, lead0:s1], lead0, f, p, W, SEED, ch, c0, 3, precision=64)
            seg = got[ch, c0 * olen:(c0 + 3) * olen]
            assert np.abs(seg - 2.0 * ref).max() <= 2 * TOL_MAX_ABS and ol.snr_db(2.0 * ref, seg) > TOL_SNR_DB
    with rb.Engine(8, rate, W, f, 1.0, p, seed=SEED) as e:   # x2 is exact in binary floating point: 2 * plain, to rounding
        plain = e.process_all(xh)
    assert np.abs(got - 2.0 * plain).max() < 1e-5
