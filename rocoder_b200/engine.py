"""Engine: thin object wrapper over the C ABI (include/rocoder_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import RcgError, WouldBlock, check, lib, make_config, rcg_plan_info, rcg_segment


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the handle that explicitly names the default stream


def _stream(stream: Optional[int]) -> Optional[int]:
    """None -> the engine's own stream; 0 (torch's default stream handle) -> cudaStreamLegacy"""
    if stream is None:
        return None
    return CUDA_STREAM_LEGACY if stream == 0 else stream


def _ptr(x) -> Optional[int]:
    """device/host pointer of a torch tensor, numpy array, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError(f"cannot take a pointer of {type(x)}")


def hanning(n: int) -> np.ndarray:
    """windows::hanning (reference windows.rs:4-9)"""
    out = np.empty(n, np.float32)
    check(lib().rcg_hanning(out.ctypes.data, n))
    return out


def hanning_crossfade_compensation(n: int) -> np.ndarray:
    """crossfade::hanning_crossfade_compensation (reference crossfade.rs:4-10)"""
    out = np.empty(n, np.float32)
    check(lib().rcg_hanning_crossfade_compensation(out.ctypes.data, n))
    return out


def phase(seed: int, channel: int, frame: int, bin_: int, window_len: int) -> float:
    return float(lib().rcg_phase(seed, channel, frame, bin_, window_len))


def _plan_dict(p: rcg_plan_info) -> dict:
    return {name: getattr(p, name) for name, _ in rcg_plan_info._fields_}


def plan(factor: float = 1.0, pitch_multiple: int = 1, window_len: int = 16384, amplitude: float = 1.0,
         channels: int = 1, sample_rate: int = 44100) -> dict:
    """Constants of Stretcher::new (reference stretcher.rs:40-59). Pure host arithmetic."""
    cfg = make_config(channels, sample_rate, window_len, factor, amplitude, pitch_multiple)
    p = rcg_plan_info()
    check(lib().rcg_plan(C.byref(cfg), C.byref(p)))
    return _plan_dict(p)


def expected_counts(n_in: int, factor: float = 1.0, pitch_multiple: int = 1, window_len: int = 16384):
    """(calls, frames, output samples) per channel (stop rule of stretcher.rs:123-135)."""
    cfg = make_config(1, 44100, window_len, factor, 1.0, pitch_multiple)
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib().rcg_expected_counts(C.byref(cfg), n_in, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def channel_bound(window_len: int, sample_rate: int, buffer_dur_s: float) -> int:
    cfg = make_config(1, sample_rate, window_len, buffer_dur_s=buffer_dur_s)
    b = C.c_uint64()
    check(lib().rcg_channel_bound(C.byref(cfg), C.byref(b)))
    return b.value


def fade_in_out(d_audio, fade_s: float, sample_rate: int, expected_total_samples: int, first_pos: int = 0,
                stream: Optional[int] = None) -> None:
    """`-x` fade of the reference's live mixer (mixer.rs:132-190) applied in place to a CUDA tensor [C, n]"""
    check(lib().rcg_fade_in_out(d_audio.data_ptr(), d_audio.stride(0), d_audio.shape[0], d_audio.shape[1], first_pos,
                                fade_s, sample_rate, expected_total_samples, _stream(stream) if stream is not None else None))


def fade_keyframes(fade_s: float, sample_rate: int, expected_total_samples: int):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib().rcg_fade_keyframes(fade_s, sample_rate, expected_total_samples, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def bind_host_thread(device: int = 0) -> int:
    """Pin the calling thread to the NUMA node of `device` (rcg_bind_host_thread); returns the node or -1."""
    node = C.c_int(-1)
    check(lib().rcg_bind_host_thread(device, C.byref(node)))
    return node.value


def device_count() -> int:
    """CUDA devices the library can run on (sm_100)"""
    n = C.c_int()
    check(lib().rcg_device_count(C.byref(n)))
    return n.value


def measure_fp32_peak(reps: int = 5) -> float:
    v = C.c_double()
    check(lib().rcg_measure_fp32_peak(C.byref(v), reps))
    return v.value


class PinnedBuffer:
    """float32 host buffer in page-locked memory (rcg_host_alloc)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        n = int(np.prod(self.shape)) if self.shape else 1
        self._ptr = C.c_void_p()
        check(lib().rcg_host_alloc(C.byref(self._ptr), max(n, 1) * 4))
        buf = (C.c_float * max(n, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float32, count=n).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib().rcg_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """All channels of one stretch job on one GPU, or -- `devices=[0, 1, ...]` -- partitioned over several
    (by channel when the channels divide evenly, else by time; rcg_config.n_devices)."""

    def __init__(self, channels: int, sample_rate: int, window_len: int = 16384, factor: float = 1.0,
                 amplitude: float = 1.0, pitch_multiple: int = 1, buffer_dur_s: float = 1.0, seed: int = 0,
                 device: int = 0, flags: int = 0, max_lookahead_calls: int = 0, first_channel: int = 0,
                 devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        self.cfg = make_config(channels, sample_rate, window_len, factor, amplitude, pitch_multiple, buffer_dur_s,
                               seed, device, flags, max_lookahead_calls, first_channel, devices)
        check(lib().rcg_create(C.byref(self.cfg), C.byref(self._h)))
        p = rcg_plan_info()
        check(lib().rcg_get_plan(self._h, C.byref(p)))
        self.plan = _plan_dict(p)
        self.channels = channels
        self._keep = []  # plugin callbacks must outlive the engine

    def close(self):
        if self._h is not None and self._h.value:
            lib().rcg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_window(self, window) -> None:
        """replace the Hann window (the `window: Vec<f32>` argument of Stretcher::new)"""
        w = np.ascontiguousarray(window, dtype=np.float32)
        check(lib().rcg_set_window(self._h, w.ctypes.data, w.size))

    # -- batched ---------------------------------------------------------------------------
    def expected_counts(self, n_in: int):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().rcg_expected_counts(C.byref(self.cfg), n_in, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def process_all_into(self, inp: np.ndarray, out: np.ndarray) -> int:
        """inp [C, N] float32 -> out [C, >= out_len] float32 (host buffers). Returns out_len."""
        assert inp.dtype == np.float32 and out.dtype == np.float32
        assert inp.ndim == 2 and out.ndim == 2 and inp.shape[0] == self.channels == out.shape[0]
        assert (inp.shape[1] == 0 or inp.strides[1] == 4) and out.strides[1] == 4
        ins = (C.c_void_p * self.channels)(*[inp[c].ctypes.data for c in range(self.channels)])
        outs = (C.c_void_p * self.channels)(*[out[c].ctypes.data for c in range(self.channels)])
        n_out = (C.c_uint64 * self.channels)()
        check(lib().rcg_process_all(self._h, ins, inp.shape[1], outs, out.shape[1], n_out))
        return int(n_out[0])

    def last_partition(self) -> dict:
        """how the last process_all was spread over the engine's devices"""
        n, ex, split = C.c_uint32(), C.c_uint32(), C.c_char_p()
        check(lib().rcg_last_partition(self._h, C.byref(n), C.byref(split), C.byref(ex)))
        return {"devices_used": n.value, "split": split.value.decode(), "exchanges": ex.value}

    def process_all(self, inp: np.ndarray) -> np.ndarray:
        inp = np.ascontiguousarray(inp, dtype=np.float32)
        if inp.ndim == 1:
            inp = inp[None, :]
        _, _, out_len = self.expected_counts(inp.shape[1])
        out = np.empty((self.channels, out_len), np.float32)
        self.process_all_into(inp, out)
        return out

    def process_segment(self, d_in, in_stride: int, in_origin: int, in_valid: int, d_out, out_stride: int,
                        first_call: int, n_calls: int, tail_in=None, tail_out=None, head_out=None,
                        defer_first: bool = False, stream: Optional[int] = None) -> None:
        s = rcg_segment()
        s.d_in = _ptr(d_in); s.in_stride = in_stride; s.in_origin = in_origin; s.in_valid = in_valid
        s.d_out = _ptr(d_out); s.out_stride = out_stride
        s.first_call = first_call; s.n_calls = n_calls
        s.d_tail_in = _ptr(tail_in); s.d_tail_out = _ptr(tail_out); s.d_head_out = _ptr(head_out)
        s.flags = _abi.RCG_SEG_DEFER_FIRST if defer_first else 0
        s.stream = _stream(stream)
        check(lib().rcg_process_segment(self._h, C.byref(s)))

    def finish_boundary(self, d_out, out_stride: int, first_call: int, head, tail_prev,
                        stream: Optional[int] = None) -> None:
        check(lib().rcg_finish_boundary(self._h, _ptr(d_out), out_stride, first_call, _ptr(head), _ptr(tail_prev),
                                        _stream(stream)))

    def synchronize(self) -> None:
        check(lib().rcg_synchronize(self._h))

    # -- streaming -------------------------------------------------------------------------
    def send(self, channel: int, chunk) -> None:
        chunk = np.ascontiguousarray(chunk, dtype=np.float32)
        check(lib().rcg_send(self._h, channel, chunk.ctypes.data, chunk.size))

    def close_input(self, channel: int) -> None:
        check(lib().rcg_close_input(self._h, channel))

    def next_window(self, channel: int) -> np.ndarray:
        out = np.empty(self.plan["out_len_per_call"], np.float32)
        n = C.c_uint64()
        check(lib().rcg_next_window(self._h, channel, out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value]

    def is_done(self, channel: int) -> bool:
        d = C.c_int()
        check(lib().rcg_is_done(self._h, channel, C.byref(d)))
        return bool(d.value)

    # -- plugins ---------------------------------------------------------------------------
    def push_frequency_kernel(self, fn) -> None:
        """fn(elapsed_ms, spectrum: complex64[W]) -> complex64[W]; raising stands for a panic."""
        W = self.cfg.window_len

        def tramp(elapsed_ms, pin, n_bins, pout):
            try:
                a = np.ctypeslib.as_array(pin, shape=(2 * n_bins,)).view(np.complex64)
                r = np.asarray(fn(int(elapsed_ms), a), dtype=np.complex64)
                if r.shape[0] != n_bins:
                    return 2
                np.ctypeslib.as_array(pout, shape=(2 * n_bins,))[:] = r.view(np.float32)
                return 0
            except Exception:
                return 1

        cb = _abi.APPLY_FN(tramp)
        self._keep.append(cb)
        check(lib().rcg_push_frequency_kernel(self._h, cb))
        del W

    def push_device_frequency_kernel(self, cuda_source: str) -> None:
        """CUDA source defining `__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins,
        const float2* spectrum)`; compiled with NVRTC for sm_100a, spectra stay in HBM."""
        check(lib().rcg_push_device_frequency_kernel(self._h, cuda_source.encode()))

    def push_local_device_frequency_kernel(self, cuda_source: str) -> None:
        """CUDA source defining the bin-local form `__device__ float2 apply(size_t elapsed_ms, unsigned bin,
        unsigned n_bins, float2 x)`; NVRTC builds the stretch kernel around it, no spectra in memory."""
        check(lib().rcg_push_local_device_frequency_kernel(self._h, cuda_source.encode()))

    def load_frequency_kernel(self, so_path: str) -> None:
        check(lib().rcg_load_frequency_kernel(self._h, so_path.encode()))

    def watch_frequency_kernel(self, path: str, poll_ms: int = 0) -> None:
        """hotswapper::hotswap (reference hotswapper.rs:14-33): build `path` (.cu / .c / .cpp / .so) now,
        rebuild it whenever it changes (polled every poll_ms, 0 = 100 ms) and swap each good build in at
        the start of the next batch; a failed build keeps the current kernel."""
        check(lib().rcg_watch_frequency_kernel(self._h, str(path).encode(), poll_ms))

    def poll_frequency_kernels(self) -> int:
        n = C.c_uint32()
        check(lib().rcg_poll_frequency_kernels(self._h, C.byref(n)))
        return n.value

    def watch_stats(self) -> dict:
        ok, bad, sw = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().rcg_watch_stats(self._h, C.byref(ok), C.byref(bad), C.byref(sw)))
        return {"builds_ok": ok.value, "builds_failed": bad.value, "swaps": sw.value}

    def watch_log(self) -> str:
        buf = C.create_string_buffer(4096)
        check(lib().rcg_watch_log(self._h, buf, len(buf)))
        return buf.value.decode(errors="replace")

    def set_clock_base(self, base_ms: int) -> None:
        """RCG_FLAG_FRAME_CLOCK: elapsed_ms of frame 0 (default: the wall clock at the first processing call)"""
        check(lib().rcg_set_clock_base(self._h, base_ms))

    def frequency_kernel_count(self) -> int:
        n = C.c_uint32()
        check(lib().rcg_frequency_kernel_count(self._h, C.byref(n)))
        return n.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib().rcg_launch_count(self._h, C.byref(n)))
        return n.value

    def launch_geometry(self) -> dict:
        """threads / frame lanes / shared memory per CTA, CTAs per SM and the largest grid of the stretch kernel"""
        t, l, c, g = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        sm = C.c_uint64()
        check(lib().rcg_launch_geometry(self._h, C.byref(t), C.byref(l), C.byref(sm), C.byref(c), C.byref(g)))
        return {"threads": t.value, "lanes_per_cta": l.value, "smem_bytes": sm.value, "ctas_per_sm": c.value, "max_grid": g.value}


__all__ = ["Engine", "PinnedBuffer", "bind_host_thread", "device_count", "fade_in_out", "fade_keyframes", "measure_fp32_peak", "RcgError", "WouldBlock", "hanning", "hanning_crossfade_compensation",
           "plan", "expected_counts", "channel_bound", "phase"]
