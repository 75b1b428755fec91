"""Synthetic inputs for tests and bench: planar float32 [channels, samples]."""
from __future__ import annotations

import numpy as np


def sine_sweep(n: int, sample_rate: int, channels: int = 2, f0: float = 20.0, f1: float = 20000.0) -> np.ndarray:
    """0.5*sin(2 pi (f0 t + (f1-f0) t^2 / (2T))) with a per-channel phase offset c*pi/4."""
    t = np.arange(n, dtype=np.float64) / sample_rate
    T = max(n / sample_rate, 1e-9)
    ph = 2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * T))
    return np.stack([0.5 * np.sin(ph + c * np.pi / 4.0) for c in range(channels)]).astype(np.float32)


def noise_chirp(n: int, sample_rate: int, channels: int = 2, seed: int = 0) -> np.ndarray:
    """0.25*U(-1,1) white noise + 0.25*sweep (i.e. half of sine_sweep)."""
    rng = np.random.default_rng(seed)
    noise = (rng.random((channels, n), dtype=np.float32) * 2.0 - 1.0) * np.float32(0.25)
    return (noise + np.float32(0.5) * sine_sweep(n, sample_rate, channels)).astype(np.float32)


# ---- position-addressable variant: sample i of channel c is a pure function of (seed, c, i) -------------------
# Any span of a long signal can be synthesised on its own -- by another rank, or for a halo / boundary check -- and
# is identical to the same span cut out of the whole.  Noise = murmur3 finaliser of the sample index; the sweep runs
# over a fixed T (total_seconds), not over the span.
_M32 = 0xFFFFFFFF


def _noise_u32(idx, channel: int, seed: int):
    """idx: int64 array (numpy) or tensor (torch) of global sample indices -> uniform 32-bit values in int64"""
    h = (idx & _M32) ^ (((idx >> 32) * 0x9E3779B1) & _M32) ^ ((seed * 0x85EBCA77 + channel * 0xC2B2AE3D + 0x27D4EB2F) & _M32)
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & _M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & _M32
    return h ^ (h >> 16)


def noise_chirp_span(s0: int, s1: int, sample_rate: int, channels: int = 2, seed: int = 0,
                     total_seconds: float = 600.0, f0: float = 20.0, f1: float = 20000.0) -> np.ndarray:
    """samples [s0, s1) of the endless noise+chirp signal: 0.25*U(-1,1) + 0.25*sin(sweep phase + c*pi/4)"""
    idx = np.arange(s0, s1, dtype=np.int64)
    t = idx.astype(np.float64) / sample_rate
    ph = 2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * total_seconds))
    out = np.empty((channels, s1 - s0), np.float32)
    with np.errstate(over="ignore"):
        for c in range(channels):
            u = _noise_u32(idx, c, seed).astype(np.float64) * (1.0 / 4294967296.0)
            out[c] = (0.25 * (2.0 * u - 1.0) + 0.25 * np.sin(ph + c * np.pi / 4.0)).astype(np.float32)
    return out


def noise_chirp_span_torch(s0: int, s1: int, sample_rate: int, channels: int, seed: int, device,
                           total_seconds: float = 600.0, f0: float = 20.0, f1: float = 20000.0, first_channel: int = 0):
    """the same signal synthesised on a CUDA device, in pieces of 2^24 samples (fp64 phase: a few hundred MB of
    temporaries at most); equal to noise_chirp_span up to the last bit of sin()"""
    import math
    import torch
    out = torch.empty((channels, s1 - s0), dtype=torch.float32, device=device)
    piece = 1 << 24
    for a in range(s0, s1, piece):
        b = min(a + piece, s1)
        idx = torch.arange(a, b, dtype=torch.int64, device=device)
        t = idx.to(torch.float64) / sample_rate
        ph = 2.0 * math.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * total_seconds))
        for c in range(channels):
            u = _noise_u32(idx, first_channel + c, seed).to(torch.float64) * (1.0 / 4294967296.0)
            out[c, a - s0:b - s0] = (0.25 * (2.0 * u - 1.0) + 0.25 * torch.sin(ph + (first_channel + c) * math.pi / 4.0)).to(torch.float32)
    return out
