"""ctypes binding of include/rocoder_b200.h (librocoder_b200.so).

The library is CUDA-only.  Loading it never falls back to anything else: if the shared
object is missing, or a GPU entry point is called without an sm_100 device, the error
is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCG_LIB selects another build of the same library (scripts/variant_bench.sh); never a fallback
LIB_PATH = os.environ.get("RCG_LIB") or os.path.join(_HERE, "librocoder_b200.so")

RCG_OK = 0
RCG_ERR_INVALID_ARG = 1
RCG_ERR_PITCH = 2
RCG_ERR_WINDOW = 3
RCG_ERR_STEP = 4
RCG_ERR_CUDA = 5
RCG_ERR_NO_DEVICE = 6
RCG_ERR_WOULD_BLOCK = 7
RCG_ERR_CAPACITY = 8
RCG_ERR_PLUGIN = 9
RCG_ERR_STATE = 10

RCG_FLAG_RANDOM_SEED = 1
RCG_FLAG_BYPASS_RESYNTH = 2
RCG_FLAG_SHARD_HALO = 4
RCG_FLAG_SHARD_ALWAYS = 8
RCG_FLAG_FRAME_CLOCK = 16
RCG_MAX_DEVICES = 16
RCG_SEG_DEFER_FIRST = 1


class RcgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"rocoder_b200: {message} (status {status})")
        self.status = status


class WouldBlock(RcgError):
    """next_window needs input that was not sent yet (recv() would block, stretcher.rs:125)."""


class rcg_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("channels", C.c_uint32), ("sample_rate", C.c_uint32),
        ("window_len", C.c_uint32), ("factor", C.c_float), ("amplitude", C.c_float),
        ("pitch_multiple", C.c_int32), ("buffer_dur_s", C.c_float), ("seed", C.c_uint64),
        ("device", C.c_int32), ("flags", C.c_uint32), ("max_lookahead_calls", C.c_uint32),
        ("first_channel", C.c_uint32),
        ("n_devices", C.c_uint32), ("devices", C.c_int32 * RCG_MAX_DEVICES),
    ]


class rcg_plan_info(C.Structure):
    _fields_ = [
        ("pitch_shifted_factor", C.c_float), ("corrected_amp_factor", C.c_float),
        ("window_len", C.c_uint64), ("half_window_len", C.c_uint64),
        ("samples_needed_per_window", C.c_uint64), ("sample_step_len", C.c_uint64),
        ("frames_per_call", C.c_uint64), ("out_len_per_call", C.c_uint64),
    ]


class rcg_segment(C.Structure):
    _fields_ = [
        ("d_in", C.c_void_p), ("in_stride", C.c_int64), ("in_origin", C.c_int64), ("in_valid", C.c_int64),
        ("d_out", C.c_void_p), ("out_stride", C.c_int64), ("first_call", C.c_uint64), ("n_calls", C.c_uint64),
        ("d_tail_in", C.c_void_p), ("d_tail_out", C.c_void_p), ("d_head_out", C.c_void_p),
        ("flags", C.c_uint32), ("reserved", C.c_uint32), ("stream", C.c_void_p),
    ]


APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float))

_lib = None


def lib() -> C.CDLL:
    """Load librocoder_b200.so; raises if it has not been built (python -m rocoder_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RcgError(RCG_ERR_STATE, f"{LIB_PATH} is missing: build it with `python -m rocoder_b200.build` "
                                      "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    sig = {
        "rcg_plan": (C.c_int, [P(rcg_config), P(rcg_plan_info)]),
        "rcg_expected_counts": (C.c_int, [P(rcg_config), C.c_uint64, P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)]),
        "rcg_hanning": (C.c_int, [C.c_void_p, C.c_size_t]),
        "rcg_hanning_crossfade_compensation": (C.c_int, [C.c_void_p, C.c_size_t]),
        "rcg_channel_bound": (C.c_int, [P(rcg_config), P(C.c_uint64)]),
        "rcg_phase": (C.c_float, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]),
        "rcg_last_error": (C.c_char_p, []),
        "rcg_status_string": (C.c_char_p, [C.c_int]),
        "rcg_abi_version": (C.c_int, []),
        "rcg_create": (C.c_int, [P(rcg_config), P(C.c_void_p)]),
        "rcg_destroy": (None, [C.c_void_p]),
        "rcg_set_window": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "rcg_get_plan": (C.c_int, [C.c_void_p, P(rcg_plan_info)]),
        "rcg_process_all": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_uint64, P(C.c_void_p), C.c_uint64, P(C.c_uint64)]),
        "rcg_last_partition": (C.c_int, [C.c_void_p, P(C.c_uint32), P(C.c_char_p), P(C.c_uint32)]),
        "rcg_process_segment": (C.c_int, [C.c_void_p, P(rcg_segment)]),
        "rcg_finish_boundary": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
        "rcg_synchronize": (C.c_int, [C.c_void_p]),
        "rcg_send": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
        "rcg_close_input": (C.c_int, [C.c_void_p, C.c_uint32]),
        "rcg_next_window": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, P(C.c_uint64)]),
        "rcg_is_done": (C.c_int, [C.c_void_p, C.c_uint32, P(C.c_int)]),
        "rcg_push_frequency_kernel": (C.c_int, [C.c_void_p, APPLY_FN]),
        "rcg_load_frequency_kernel": (C.c_int, [C.c_void_p, C.c_char_p]),
        "rcg_push_device_frequency_kernel": (C.c_int, [C.c_void_p, C.c_char_p]),
        "rcg_push_local_device_frequency_kernel": (C.c_int, [C.c_void_p, C.c_char_p]),
        "rcg_frequency_kernel_count": (C.c_int, [C.c_void_p, P(C.c_uint32)]),
        "rcg_set_clock_base": (C.c_int, [C.c_void_p, C.c_uint64]),
        "rcg_watch_frequency_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32]),
        "rcg_poll_frequency_kernels": (C.c_int, [C.c_void_p, P(C.c_uint32)]),
        "rcg_watch_stats": (C.c_int, [C.c_void_p, P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]),
        "rcg_watch_log": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
        "rcg_fade_in_out": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_float, C.c_uint32,
                                      C.c_uint64, C.c_void_p]),
        "rcg_fade_keyframes": (C.c_int, [C.c_float, C.c_uint32, C.c_uint64, P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)]),
        "rcg_measure_fp32_peak": (C.c_int, [P(C.c_double), C.c_int]),
        "rcg_device_count": (C.c_int, [P(C.c_int)]),
        "rcg_host_alloc": (C.c_int, [P(C.c_void_p), C.c_size_t]),
        "rcg_bind_host_thread": (C.c_int, [C.c_int, P(C.c_int)]),
        "rcg_host_free": (C.c_int, [C.c_void_p]),
        "rcg_launch_count": (C.c_int, [C.c_void_p, P(C.c_uint64)]),
        "rcg_launch_geometry": (C.c_int, [C.c_void_p, P(C.c_uint32), P(C.c_uint32), P(C.c_uint64), P(C.c_uint32), P(C.c_uint32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status: int) -> None:
    if status == RCG_OK:
        return
    msg = lib().rcg_last_error().decode("utf-8", "replace")
    if status == RCG_ERR_WOULD_BLOCK:
        raise WouldBlock(status, msg)
    raise RcgError(status, msg)


def make_config(channels: int, sample_rate: int, window_len: int = 16384, factor: float = 1.0,
                amplitude: float = 1.0, pitch_multiple: int = 1, buffer_dur_s: float = 1.0, seed: int = 0,
                device: int = 0, flags: int = 0, max_lookahead_calls: int = 0, first_channel: int = 0,
                devices=None) -> rcg_config:
    cfg = rcg_config()
    cfg.struct_size = C.sizeof(rcg_config)
    cfg.channels = channels
    cfg.sample_rate = sample_rate
    cfg.window_len = window_len
    cfg.factor = factor
    cfg.amplitude = amplitude
    cfg.pitch_multiple = pitch_multiple
    cfg.buffer_dur_s = buffer_dur_s
    cfg.seed = seed & 0xFFFFFFFFFFFFFFFF
    cfg.device = device
    cfg.flags = flags
    cfg.max_lookahead_calls = max_lookahead_calls
    cfg.first_channel = first_channel
    if devices is not None:   # several GPUs behind one engine (the same ordinal may repeat)
        devices = [int(d) for d in devices]
        if not 1 <= len(devices) <= RCG_MAX_DEVICES:
            raise ValueError(f"1 to {RCG_MAX_DEVICES} devices")
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        cfg.device = devices[0]
    return cfg
