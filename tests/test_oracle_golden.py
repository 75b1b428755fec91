"""The oracle (and the product's [host] arithmetic) against every golden vector the
reference's own tests hold for this path."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import rocoder_b200 as rb

EPS = 1.0e-4  # test_utils.rs:4

# windows.rs:32-41
HANNING_32 = [
    0.0, 0.010235041, 0.040521085, 0.089618266, 0.15551656, 0.23551801, 0.32634735, 0.4242862,
    0.5253246, 0.6253263, 0.7201971, 0.80605304, 0.87937903, 0.93717337, 0.9770697, 0.9974347,
    0.9974346, 0.9770696, 0.9371733, 0.87937903, 0.8060529, 0.720197, 0.62532616, 0.5253247,
    0.4242862, 0.32634744, 0.23551783, 0.15551639, 0.08961815, 0.040521085, 0.010235012, 0.0,
]


def test_hanning_result():
    """windows.rs:30-43 -- here the restatement reproduces the literals bit for bit."""
    want = np.array(HANNING_32, np.float32)
    assert np.abs(ol.hanning(32) - want).max() < EPS
    assert np.array_equal(ol.hanning(32), want)
    assert np.array_equal(rb.hanning(32), want)


def test_inv_hanning():
    """windows.rs:52-57"""
    L = ol.oracle()
    src = np.array([1.0, 0.7, 0.3], np.float32)
    out = np.zeros(3, np.float32)
    L.ro_inverse(src.ctypes.data_as(ol.PF), out.ctypes.data_as(ol.PF), C.c_size_t(3))
    assert np.abs(out - np.array([1.0, 1.4285715, 3.3333333], np.float32)).max() < EPS


def test_rectangular_result():
    """windows.rs:46-49"""
    out = np.zeros(4, np.float32)
    ol.oracle().ro_rectangular(out.ctypes.data_as(ol.PF), C.c_size_t(4))
    assert np.array_equal(out, np.ones(4, np.float32))


def test_resample_noop():
    """resampler.rs:43-46"""
    v = np.array([1.0, 2.0, 3.0], np.float32)
    assert np.array_equal(ol.resample(v, 1), v)


def test_resample_faster():
    """resampler.rs:49-54"""
    assert np.array_equal(ol.resample(np.array([1, 2, 3, 4], np.float32), 2), np.array([1, 3], np.float32))


def test_resample_slower_literal():
    """resampler.rs:20-35 has no reference test; pin the literal reading: (len-1)*q outputs."""
    out = ol.resample(np.array([0.0, 3.0, 6.0], np.float32), -3)
    assert np.array_equal(out, np.array([0, 1, 2, 3, 4, 5], np.float32))


def test_resample_invalid():
    """resampler.rs:11 panics on 0 and -1"""
    for p in (0, -1):
        with pytest.raises(ValueError):
            ol.resample(np.ones(4, np.float32), p)


@pytest.mark.parametrize("a,b,r,want", [(0.0, 1.0, 0.6, 0.6), (234.287, 239847.45, 0.6, 144002.18480000002),
                                        (10.0, -20.0, 0.5, -5.0)])
def test_lerp(a, b, r, want):
    """math.rs:68-80"""
    got = ol.oracle().ro_lerp(a, b, r)
    assert abs(got - want) < EPS * max(1.0, abs(want))


def test_sqrt_interp_fade_samples():
    """audio.rs:290-313 expects fade-in gains [0.5, 0.70710677, 0.8660254] at p/dur = 1/4, 2/4, 3/4"""
    L = ol.oracle()
    got = [L.ro_sqrt_interp(0.0, 1.0, r) for r in (0.25, 0.5, 0.75)]
    assert np.allclose(got, [0.5, 0.70710677, 0.8660254], atol=EPS)


def _basic_stretcher(window_len):
    # stretcher.rs:163-179: rectangular window, factor 1, amplitude 1, pitch 1, 44.1 kHz stereo
    return ol.OracleStretcher(1.0, 1, window_len, window=np.ones(window_len, np.float32), sample_rate=44100,
                              channels=2, buffer_dur=1.0)


def test_ensure_input_samples_available_when_channel_closed_fills_with_zeros():
    """stretcher.rs:145-151"""
    s = _basic_stretcher(1000)
    s.close()
    assert s.ensure_input_samples_available(4) == 0
    assert s.is_done()
    assert np.array_equal(s.input_buf(), np.zeros(4, np.float32))


def test_ensure_input_samples_available_loading_multiple_chunks():
    """stretcher.rs:153-161"""
    s = _basic_stretcher(1000)
    s.send([1.0, 2.0, 3.0])
    s.send([4.0, 5.0])
    assert s.ensure_input_samples_available(4) == 0
    assert not s.is_done()
    assert np.array_equal(s.input_buf(), np.array([1, 2, 3, 4, 5], np.float32))


def test_non_power_of_two_window_literal_loop():
    """The oracle keeps the reference's generality (README.md:100): W=1000 runs through the DFT path."""
    x = np.random.default_rng(0).standard_normal((1, 3000)).astype(np.float32) * 0.1
    s = ol.OracleStretcher(1.0, 1, 1000)
    s.set_philox(1, 0)
    s.send(x[0])
    s.close()
    n = 0
    while not s.is_done():
        w = s.next_window()
        assert w.size == 1000 and np.isfinite(w).all()
        n += 1
    assert n == 3  # j1 = (3000-1000)/500 + 1 = 5 -> calls = 5 // 2 + 1
