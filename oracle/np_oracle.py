"""np_oracle.py -- second, independent restatement of rocoder's stretch path in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/rocoder_oracle.h).  Where the C oracle restates
the reference literally (deque loop, own Stockham FFT), this one uses the CLOSED FORM
of the indexing and scipy's pocketfft, so that the two restatements
check each other: frame j reads in[j*step : j*step+W) zero padded, output block m is
(y_m[:H] + y_{m-1}[H:]) * env * amp, call c emits resample(z[c*k*H : c*k*H+S], p).

Reference lines followed (relative to /root/reference/src):
  windows.rs:4-9, crossfade.rs:4-10, stretcher.rs:40-59 + 87-135, fft.rs:13 + 42-74,
  resampler.rs:3-35, math.rs:28-30, stretcher_processor.rs:56-71.
"""
from __future__ import annotations

import numpy as np

try:  # pocketfft with native float32 support
    from scipy import fft as _fft
except Exception:  # pragma: no cover
    _fft = None

F32 = np.float32
PI_F32 = F32(np.pi)


def hanning(n: int) -> np.ndarray:
    """windows.rs:4-9 (f32 arithmetic; cos evaluated by numpy's f32 loop)."""
    i = np.arange(n, dtype=F32)
    two_pi = F32(PI_F32 * F32(2.0))
    return (F32(0.5) - np.cos((i * two_pi) / F32(n - 1), dtype=F32) * F32(0.5)).astype(F32)


def hanning_crossfade_compensation(n: int) -> np.ndarray:
    """crossfade.rs:4-10."""
    i = np.arange(n, dtype=F32)
    two_pi = F32(PI_F32 * F32(2.0))
    hinv = F32((F32(1.0) + np.sqrt(np.sqrt(F32(0.5)))) * F32(0.5))
    return (F32(0.5) - (F32(1.0) - hinv) * np.cos((i * two_pi) / F32(n - 1), dtype=F32)).astype(F32)


def constants(factor: float, pitch_multiple: int, window_len: int, amplitude: float = 1.0):
    """stretcher.rs:40-59 -- all in f32, truncating casts."""
    p = int(pitch_multiple)
    if p == 0 or p == -1 or not -127 <= p <= 127:
        raise ValueError("invalid pitch multiple")
    f = F32(factor)
    ap = F32(abs(p))
    psf = F32(f / ap) if p < 0 else F32(f * ap)
    S = int(np.ceil(F32(window_len) / ap)) if p < 0 else window_len * abs(p)
    H = window_len // 2
    step = int(F32(window_len) / F32(psf * F32(2.0)))
    amp = F32(max(F32(4.0), F32(psf / F32(4.0))) * F32(amplitude))
    k = -(-S // H)
    if p == 1:
        out_len = S
    elif p > 1:
        out_len = -(-S // p)
    else:
        out_len = (S - 1) * abs(p)
    return dict(psf=psf, S=S, H=H, step=step, amp=amp, k=k, out_len=out_len)


def counts(n_in: int, factor: float, pitch_multiple: int, window_len: int):
    c = constants(factor, pitch_multiple, window_len)
    if not 1 <= c["step"] <= window_len:
        raise ValueError("step outside 1..=W")
    j1 = 0 if n_in < window_len else (n_in - window_len) // c["step"] + 1
    calls = j1 // c["k"] + 1
    return calls, calls * c["k"], calls * c["out_len"]


# ---- Philox4x32-10 phase source (vectorised) ------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _M0 * c0.astype(np.uint64)
        p1 = _M1 * c2.astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ np.uint32(k0)
        n1 = (p1 & mask).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ np.uint32(k1)
        n3 = (p0 & mask).astype(np.uint32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def phase_from_u32(u: np.ndarray) -> np.ndarray:
    """rand 0.8.5 gen_range(0.0..PI) for f32 with fft.rs:13's TWO_PI == PI."""
    v = (u >> np.uint32(9)).astype(F32) * F32(2.0 ** -23)
    return (v * PI_F32 + F32(0.0)).astype(F32)


def phases_for_frame(seed: int, channel: int, frame: int, window_len: int) -> np.ndarray:
    """which Philox block / word serves a bin: the rule of rocoder_oracle.c ro_phase_slot, for every even W
    (H = W/2 may be odd: the quadruple {k, W-k, H-k, H+k} then has no self-paired middle)"""
    W, H = window_len, window_len // 2
    Q = H // 2
    b = np.arange(W)
    pair = np.where(b <= Q, b, np.where(b <= H, H - b, np.where(b < H + Q, b - H, W - b)))
    word = np.where(b <= Q, 0, np.where(b <= H, 2, np.where(b < H + Q, 3, 1)))
    blocks = np.arange(int(pair.max()) + 1, dtype=np.uint32)
    o = philox4x32_10(blocks, np.uint32(frame & 0xFFFFFFFF), np.uint32(frame >> 32),
                      np.uint32(channel), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ph = np.stack([phase_from_u32(x) for x in o])
    return ph[word, pair].astype(F32)


# ---- ReFFT (fft.rs:42-74) -------------------------------------------------------------
def resynth(samples: np.ndarray, window: np.ndarray, phases: np.ndarray, dtype=F32,
            plugin=None) -> np.ndarray:
    W = window.shape[0]
    cdt = np.complex64 if dtype == F32 else np.complex128
    buf = (samples.astype(dtype) * window.astype(dtype)).astype(cdt)
    spec = _fft.fft(buf)                                    # fft.rs:59
    if plugin is not None:
        spec = plugin(spec).astype(cdt)
    ph = phases.astype(dtype)
    mag = np.hypot(spec.real, spec.imag).astype(dtype)      # c.norm()
    syn = ((np.cos(ph) * mag) + 1j * (np.sin(ph) * mag)).astype(cdt)
    inv = _fft.ifft(syn, norm="forward")                    # unnormalised inverse, fft.rs:69
    return ((inv.real.astype(dtype) / dtype(W)) * window.astype(dtype)).astype(dtype)


def resample(z: np.ndarray, p: int) -> np.ndarray:
    """resampler.rs:3-35."""
    if p == 1:
        return z.copy()
    if p > 1:
        return z[::p].copy()
    if p < -1:
        q = -p
        cur, nxt = z[:-1, None], z[1:, None]
        ratio = (np.arange(q, dtype=F32) / F32(q)).astype(z.dtype)[None, :]
        return (cur + (nxt - cur) * ratio).reshape(-1).astype(z.dtype)
    raise ValueError("invalid resample factor")


def stretch_channel(x: np.ndarray, factor: float, pitch_multiple: int, window_len: int,
                    amplitude: float = 1.0, seed: int = 0, channel: int = 0, dtype=F32,
                    phase_mode: str = "philox", plugin=None, first_call: int = 0,
                    n_calls: int | None = None) -> np.ndarray:
    """Closed-form restatement of Stretcher::next_window driven to completion."""
    W = window_len
    c = constants(factor, pitch_multiple, W, amplitude)
    H, S, step, k, amp = c["H"], c["S"], c["step"], c["k"], c["amp"]
    calls, _, _ = counts(len(x), factor, pitch_multiple, W)
    last_call = calls if n_calls is None else min(calls, first_call + n_calls)
    window = hanning(W)
    env = hanning_crossfade_compensation(H).astype(dtype)
    xp = np.concatenate([x.astype(F32), np.zeros(calls * k * step + W, F32)])

    def frame(j: int) -> np.ndarray:
        s = xp[j * step: j * step + W]
        if phase_mode == "bypass":
            return ((s.astype(dtype) * window.astype(dtype)) * window.astype(dtype)).astype(dtype)
        return resynth(s, window, phases_for_frame(seed, channel, j, W), dtype, plugin)

    out = []
    prev_tail = np.zeros(H, dtype) if first_call == 0 else frame(first_call * k - 1)[H:]
    for call in range(first_call, last_call):
        z = np.empty(k * H, dtype)
        for r in range(k):
            y = frame(call * k + r)
            z[r * H:(r + 1) * H] = ((y[:H] + prev_tail) * env) * dtype(amp)
            prev_tail = y[H:]
        out.append(resample(z[:S], pitch_multiple))
    return np.concatenate(out) if out else np.zeros(0, dtype)
