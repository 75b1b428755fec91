"""ctypes access to the parity oracle (oracle/) and the host emulator (tests/emul/).
TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PF = C.POINTER(C.c_float)
PD = C.POINTER(C.c_double)

RO_PHASE_PHILOX, RO_PHASE_TABLE, RO_PHASE_BYPASS = 0, 1, 2
APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, PF, C.c_size_t, PF)


class ro_constants(C.Structure):
    _fields_ = [("pitch_shifted_factor", C.c_float), ("corrected_amp_factor", C.c_float),
                ("window_len", C.c_size_t), ("half_window_len", C.c_size_t),
                ("samples_needed_per_window", C.c_size_t), ("sample_step_len", C.c_size_t),
                ("frames_per_call", C.c_size_t), ("out_len_per_call", C.c_size_t)]


_oracle = None
_emul = None


def oracle(path: str | None = None) -> C.CDLL:
    global _oracle
    if _oracle is not None and path is None:
        return _oracle
    from rocoder_b200 import build
    L = C.CDLL(path or build.build_oracle())
    L.ro_lerp.restype = C.c_float
    L.ro_lerp.argtypes = [C.c_float] * 3
    L.ro_sqrt_interp.restype = C.c_float
    L.ro_sqrt_interp.argtypes = [C.c_float] * 3
    L.ro_phase.restype = C.c_float
    L.ro_phase.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
    L.ro_phase_from_u32.restype = C.c_float
    L.ro_phase_from_u32.argtypes = [C.c_uint32]
    L.ro_resample_len.restype = C.c_size_t
    L.ro_resample_len.argtypes = [C.c_size_t, C.c_int]
    L.ro_resample.argtypes = [PF, C.c_size_t, C.c_int, PF, C.POINTER(C.c_size_t)]
    L.ro_stretcher_new.restype = C.c_void_p
    L.ro_stretcher_new.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_int, PF, C.c_size_t,
                                   C.c_float, C.c_int]
    for name in ("ro_stretcher_free", "ro_stretcher_close", "ro_stretcher_set_phase_bypass"):
        getattr(L, name).argtypes = [C.c_void_p]
        getattr(L, name).restype = None
    L.ro_stretcher_set_phase_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    L.ro_stretcher_set_phase_table.argtypes = [C.c_void_p, PF, C.c_size_t]
    L.ro_stretcher_set_first_frame.argtypes = [C.c_void_p, C.c_uint64]
    L.ro_stretcher_push_plugin.argtypes = [C.c_void_p, APPLY_FN]
    L.ro_stretcher_send.argtypes = [C.c_void_p, PF, C.c_size_t]
    L.ro_stretcher_is_done.argtypes = [C.c_void_p]
    L.ro_stretcher_channel_bound.argtypes = [C.c_void_p]
    L.ro_stretcher_channel_bound.restype = C.c_size_t
    L.ro_stretcher_ensure_input_samples_available.argtypes = [C.c_void_p, C.c_size_t]
    L.ro_stretcher_input_buf_len.argtypes = [C.c_void_p]
    L.ro_stretcher_input_buf_len.restype = C.c_size_t
    L.ro_stretcher_input_buf_copy.argtypes = [C.c_void_p, PF]
    L.ro_stretcher_next_window.argtypes = [C.c_void_p, PF, C.POINTER(C.c_size_t)]
    L.ro_stretcher_next_window_f64.argtypes = [C.c_void_p, PD, C.POINTER(C.c_size_t)]
    L.ro_stretcher_out_len.argtypes = [C.c_void_p]
    L.ro_stretcher_out_len.restype = C.c_size_t
    L.ro_stretcher_constants.argtypes = [C.c_void_p, C.POINTER(ro_constants)]
    L.ro_stretcher_frames_done.argtypes = [C.c_void_p]
    L.ro_stretcher_frames_done.restype = C.c_uint64
    L.ro_expected_counts.argtypes = [C.c_size_t, C.c_float, C.c_int, C.c_size_t] + [C.POINTER(C.c_size_t)] * 3
    common = [C.c_uint32, C.c_uint32, C.POINTER(PF), C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_size_t,
              C.c_uint64, C.c_int, APPLY_FN]
    L.ro_process_all.restype = C.c_long
    L.ro_process_all.argtypes = common + [C.POINTER(PF), C.c_size_t, C.POINTER(C.c_size_t)]
    L.ro_process_all_f64.restype = C.c_long
    L.ro_process_all_f64.argtypes = common + [C.POINTER(PD), C.c_size_t, C.POINTER(C.c_size_t)]
    L.ro_bench_channel.restype = C.c_long
    L.ro_bench_channel.argtypes = [PF, C.c_size_t, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_size_t,
                                   C.c_uint64, C.c_uint32, C.c_size_t, PF, C.POINTER(C.c_size_t)]
    L.ro_fade_gain.restype = C.c_float
    L.ro_fade_gain.argtypes = [C.c_uint64, C.c_float, C.c_uint32, C.c_uint64]
    L.ro_fade_keyframes.argtypes = [C.c_float, C.c_uint32, C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
    if path is None:
        _oracle = L
    return L


def expected_counts(n_in: int, factor: float, pitch: int, W: int):
    a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = oracle().ro_expected_counts(n_in, factor, pitch, W, C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise ValueError(f"oracle rejects factor={factor} pitch={pitch} W={W}: {rc}")
    return a.value, b.value, c.value


def constants(factor: float, pitch: int, W: int, amplitude: float = 1.0) -> dict:
    L = oracle()
    w = np.ones(W, np.float32)
    s = L.ro_stretcher_new(1, 44100, factor, amplitude, pitch, w.ctypes.data_as(PF), W, 1.0, 32)
    if not s:
        raise ValueError("oracle rejects the pitch multiple")
    c = ro_constants()
    L.ro_stretcher_constants(s, C.byref(c))
    L.ro_stretcher_free(s)
    return {n: getattr(c, n) for n, _ in ro_constants._fields_}


def process_all(x: np.ndarray, factor: float, pitch: int, W: int, seed: int = 0, amplitude: float = 1.0,
                mode: int = RO_PHASE_PHILOX, precision: int = 32, plugin=None, sample_rate: int = 48000,
                lib: C.CDLL | None = None) -> np.ndarray:
    """The reference's file-mode run (main.rs:124-160) restated: [C, N] -> [C, out_len]."""
    L = lib or oracle()
    x = np.ascontiguousarray(x, np.float32)
    ch, n = x.shape
    calls, _, out_len = expected_counts(n, factor, pitch, W)
    dt = np.float32 if precision == 32 else np.float64
    out = np.zeros((ch, out_len), dt)
    ins = (PF * ch)(*[x[c].ctypes.data_as(PF) for c in range(ch)])
    PT = PF if precision == 32 else PD
    outs = (PT * ch)(*[out[c].ctypes.data_as(PT) for c in range(ch)])
    n_out = (C.c_size_t * ch)()
    fn = L.ro_process_all if precision == 32 else L.ro_process_all_f64
    cb = APPLY_FN(plugin) if plugin is not None else C.cast(None, APPLY_FN)
    rc = fn(ch, sample_rate, ins, n, factor, amplitude, pitch, W, seed, mode, cb, outs, out_len, n_out)
    if rc != calls:
        raise RuntimeError(f"oracle returned {rc}, expected {calls} calls")
    assert all(n_out[c] == out_len for c in range(ch))
    return out


class OracleStretcher:
    """One reference Stretcher (stretcher.rs) restated; explicit send/close like a crossbeam channel."""

    def __init__(self, factor, pitch, W, amplitude=1.0, sample_rate=44100, channels=2, window=None,
                 buffer_dur=1.0, precision=32):
        self.L = oracle()
        w = np.ascontiguousarray(window if window is not None else hanning(W), np.float32)
        self.h = self.L.ro_stretcher_new(channels, sample_rate, factor, amplitude, pitch, w.ctypes.data_as(PF), len(w),
                                         buffer_dur, precision)
        if not self.h:
            raise ValueError("invalid pitch multiple")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ro_stretcher_free(self.h)
            self.h = None

    def set_philox(self, seed, channel):
        self.L.ro_stretcher_set_phase_philox(self.h, seed, channel)

    def set_bypass(self):
        self.L.ro_stretcher_set_phase_bypass(self.h)

    def send(self, chunk):
        chunk = np.ascontiguousarray(chunk, np.float32)
        self.L.ro_stretcher_send(self.h, chunk.ctypes.data_as(PF), chunk.size)

    def close(self):
        self.L.ro_stretcher_close(self.h)

    def is_done(self):
        return bool(self.L.ro_stretcher_is_done(self.h))

    def channel_bound(self):
        return self.L.ro_stretcher_channel_bound(self.h)

    def ensure_input_samples_available(self, n):
        return self.L.ro_stretcher_ensure_input_samples_available(self.h, n)

    def input_buf(self):
        n = self.L.ro_stretcher_input_buf_len(self.h)
        out = np.zeros(n, np.float32)
        self.L.ro_stretcher_input_buf_copy(self.h, out.ctypes.data_as(PF))
        return out

    def next_window(self):
        n = self.L.ro_stretcher_out_len(self.h)
        out = np.zeros(n, np.float32)
        m = C.c_size_t()
        rc = self.L.ro_stretcher_next_window(self.h, out.ctypes.data_as(PF), C.byref(m))
        if rc:
            raise RuntimeError(f"oracle next_window: {rc}")
        return out[: m.value]


def hanning(n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    oracle().ro_hanning(out.ctypes.data_as(PF), C.c_size_t(n))
    return out


def crossfade(n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    oracle().ro_hanning_crossfade_compensation(out.ctypes.data_as(PF), C.c_size_t(n))
    return out


def resample(x: np.ndarray, p: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(max(oracle().ro_resample_len(x.size, p), 1), np.float32)
    n = C.c_size_t()
    rc = oracle().ro_resample(x.ctypes.data_as(PF), x.size, p, out.ctypes.data_as(PF), C.byref(n))
    if rc:
        raise ValueError("invalid resample factor")
    return out[: n.value]


# ---- host emulator of the kernel phases -------------------------------------------------------
_emul_local = None


def _declare_emulator(E):
    E.rcg_emul_stretch.argtypes = [C.c_int, PF, C.c_longlong, C.c_longlong, C.c_longlong, PF, C.c_longlong,
                                   C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_float, C.c_ulonglong, C.c_uint, C.c_int, PF, PF, PF]
    E.rcg_emul_boundary.argtypes = [C.c_int, PF, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                    C.c_int, C.c_float, PF, PF]
    E.rcg_emul_stretch_any.argtypes = E.rcg_emul_stretch.argtypes
    E.rcg_emul_boundary_any.argtypes = E.rcg_emul_boundary.argtypes
    E.rcg_emul_stretch_plugin.argtypes = [C.c_int, PF, C.c_longlong, C.c_longlong, PF, C.c_longlong,
                                          C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_ulonglong, APPLY_FN]
    return E


def emulator() -> C.CDLL:
    global _emul
    if _emul is None:
        from rocoder_b200 import build
        _emul = _declare_emulator(C.CDLL(build.build_emulator()))
    return _emul


def emulator_local_plugin() -> C.CDLL:
    """the emulator built with the fused bin-local frequency-kernel hook (`tilt` compiled in)"""
    global _emul_local
    if _emul_local is None:
        from rocoder_b200 import build
        _emul_local = _declare_emulator(C.CDLL(build.build_emulator_local_plugin()))
    return _emul_local


def _pf(a):
    return a.ctypes.data_as(PF) if a is not None else C.cast(None, PF)


def emul_segment(x: np.ndarray, in_origin: int, out: np.ndarray, factor: float, pitch: int, W: int,
                 first_call: int, n_calls: int, seed: int = 0, amplitude: float = 1.0, run_len: int = 5,
                 bypass: bool = False, tail_in=None, tail_out=None, head_out=None, defer_first: bool = False,
                 first_channel: int = 0, lib=None):
    """Run the kernel phases on the host for calls [first_call, first_call+n_calls)."""
    c = constants(factor, pitch, W, amplitude)
    k, olen = c["frames_per_call"], c["out_len_per_call"]
    ch, n = x.shape
    flags = (1 if bypass else 0) | (2 if defer_first else 0)
    pow2 = W & (W - 1) == 0
    E = lib or emulator()
    rc = (E.rcg_emul_stretch if pow2 else E.rcg_emul_stretch_any)(
        int(np.log2(W)) if pow2 else W, _pf(x), x.strides[0] // 4, in_origin, n, _pf(out), out.strides[0] // 4,
        first_call * olen, first_call * k, (first_call + n_calls) * k, run_len, ch, c["sample_step_len"], pitch,
        c["samples_needed_per_window"], c["corrected_amp_factor"], seed, flags, first_channel, _pf(tail_in),
        _pf(tail_out), _pf(head_out))
    assert rc == 0


def emul_boundary(out: np.ndarray, factor: float, pitch: int, W: int, first_call: int, head: np.ndarray,
                  tail_prev: np.ndarray, amplitude: float = 1.0):
    c = constants(factor, pitch, W, amplitude)
    pow2 = W & (W - 1) == 0
    rc = (emulator().rcg_emul_boundary if pow2 else emulator().rcg_emul_boundary_any)(
        int(np.log2(W)) if pow2 else W, _pf(out), out.strides[0] // 4,
                                      first_call * c["out_len_per_call"], first_call * c["frames_per_call"],
                                      out.shape[0], pitch, c["samples_needed_per_window"],
                                      c["corrected_amp_factor"], _pf(head), _pf(tail_prev))
    assert rc == 0


def emul_process_all(x: np.ndarray, factor: float, pitch: int, W: int, seed: int = 0, amplitude: float = 1.0,
                     run_len: int = 5, bypass: bool = False, lib=None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    calls, _, out_len = expected_counts(x.shape[1], factor, pitch, W)
    out = np.full((x.shape[0], out_len), np.nan, np.float32)
    emul_segment(x, 0, out, factor, pitch, W, 0, calls, seed, amplitude, run_len, bypass, lib=lib)
    return out


def emul_process_all_plugin(x: np.ndarray, factor: float, pitch: int, W: int, plugin, seed: int = 0,
                            amplitude: float = 1.0, run_len: int = 5) -> np.ndarray:
    """kernel phases of the frequency-kernel path on the host; plugin is a C-ABI `apply` callable"""
    x = np.ascontiguousarray(x, np.float32)
    c = constants(factor, pitch, W, amplitude)
    calls, frames, out_len = expected_counts(x.shape[1], factor, pitch, W)
    out = np.full((x.shape[0], out_len), np.nan, np.float32)
    cb = APPLY_FN(plugin)
    rc = emulator().rcg_emul_stretch_plugin(int(np.log2(W)), _pf(x), x.strides[0] // 4, x.shape[1], _pf(out),
                                            out.strides[0] // 4, frames, run_len, x.shape[0], c["sample_step_len"],
                                            pitch, c["samples_needed_per_window"], c["corrected_amp_factor"], seed, cb)
    assert rc == 0
    return out


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = ref.astype(np.float64)
    err = test.astype(np.float64) - ref
    den = float((err ** 2).sum())
    num = float((ref ** 2).sum())
    if den == 0.0:
        return float("inf")
    if num == 0.0:
        return float("-inf")
    return 10.0 * np.log10(num / den)


def process_calls(x_span: np.ndarray, span_origin: int, factor: float, pitch: int, W: int, seed: int, channel: int,
                  first_call: int, n_calls: int, precision: int = 64, amplitude: float = 1.0,
                  bypass: bool = False) -> np.ndarray:
    """Calls [first_call, first_call + n_calls) of ONE channel out of the middle of a long job, from the oracle.
    x_span holds the channel's samples from span_origin on and must start at or before the first sample of call
    first_call - 1: the stretcher is started one call early (phase counters set to that frame,
    ro_stretcher_set_first_frame) so that the half window carried into call first_call is the real one, and the
    extra call is discarded.  The input must reach the end of the last frame (no zero padding is wanted here)."""
    L = oracle()
    c = constants(factor, pitch, W, amplitude)
    k, step, olen = c["frames_per_call"], c["sample_step_len"], c["out_len_per_call"]
    lead = 1 if first_call > 0 else 0
    start_call = first_call - lead
    s_first = start_call * k * step
    s_end = ((first_call + n_calls) * k - 1) * step + W
    assert span_origin <= s_first and span_origin + x_span.size >= s_end, "span does not cover the calls"
    w = hanning(W)
    h = L.ro_stretcher_new(1, 48000, factor, amplitude, pitch, w.ctypes.data_as(PF), W, 1.0, precision)
    assert h
    try:
        if bypass:
            L.ro_stretcher_set_phase_bypass(h)
        else:
            L.ro_stretcher_set_phase_philox(h, seed, channel)
        L.ro_stretcher_set_first_frame(h, start_call * k)
        chunk = np.ascontiguousarray(x_span[s_first - span_origin:s_end - span_origin], np.float32)
        L.ro_stretcher_send(h, chunk.ctypes.data_as(PF), chunk.size)
        outs = []
        m = C.c_size_t()
        for _ in range(lead + n_calls):
            if precision == 64:
                o = np.zeros(olen, np.float64)
                rc = L.ro_stretcher_next_window_f64(h, o.ctypes.data_as(PD), C.byref(m))
            else:
                o = np.zeros(olen, np.float32)
                rc = L.ro_stretcher_next_window(h, o.ctypes.data_as(PF), C.byref(m))
            assert rc == 0 and m.value == olen
            outs.append(o)
        return np.concatenate(outs[lead:])
    finally:
        L.ro_stretcher_free(h)
