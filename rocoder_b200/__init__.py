"""rocoder_b200 -- B200-native engine for ajyoon/rocoder's stretch path.

Only what the path needs lives here: csrc/ (sm_100a kernels + the C ABI behind
include/rocoder_b200.h) and the host-side mirror of the reference's Stretcher /
StretcherProcessor interface.  There is no CPU fallback: without librocoder_b200.so or
without an sm_100 GPU the engine raises.
"""
from ._abi import (RCG_FLAG_BYPASS_RESYNTH, RCG_FLAG_RANDOM_SEED, RCG_FLAG_FRAME_CLOCK, RCG_FLAG_SHARD_ALWAYS, RCG_FLAG_SHARD_HALO, RcgError,
                   WouldBlock, LIB_PATH)
from .engine import (Engine, PinnedBuffer, bind_host_thread, channel_bound, device_count, expected_counts, fade_in_out, fade_keyframes, hanning, measure_fp32_peak,
                     hanning_crossfade_compensation, phase, plan)
from .stretcher import (AudioBus, AudioSpec, Disconnected, Receiver, Sender, Stretcher, StretcherProcessor,
                        bounded, stretch, unbounded)

__all__ = [
    "Engine", "PinnedBuffer", "device_count", "fade_in_out", "fade_keyframes", "measure_fp32_peak", "RcgError", "WouldBlock", "LIB_PATH", "RCG_FLAG_BYPASS_RESYNTH",
    "RCG_FLAG_RANDOM_SEED", "RCG_FLAG_FRAME_CLOCK", "RCG_FLAG_SHARD_ALWAYS", "RCG_FLAG_SHARD_HALO", "channel_bound", "expected_counts", "hanning", "hanning_crossfade_compensation",
    "phase", "plan", "AudioBus", "AudioSpec", "Disconnected", "Receiver", "Sender", "Stretcher",
    "StretcherProcessor", "bounded", "stretch", "unbounded",
]
