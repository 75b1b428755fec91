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
