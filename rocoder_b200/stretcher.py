"""Host-side mirror of the reference's signal-flow API for the stretch path.

Same names, argument meaning and error behaviour as reference src/stretcher.rs and
src/stretcher_processor.rs; the arithmetic runs in librocoder_b200.so (CUDA only).

  Receiver / Sender  ~ crossbeam_channel::unbounded (main.rs:137)
  Stretcher          ~ stretcher.rs:12-136   (one per channel)
  StretcherProcessor ~ stretcher_processor.rs:22-75
  AudioBus           ~ audio.rs:142-172
"""
from __future__ import annotations

import math
import queue
import threading
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .engine import Engine, WouldBlock, channel_bound


@dataclass(frozen=True)
class AudioSpec:
    """audio.rs:32-37"""
    channels: int
    sample_rate: int


class Disconnected(Exception):
    """RecvError: every Sender has been dropped and the channel is empty."""


class _Chan:
    def __init__(self, bound: int = 0):
        self.q: "queue.Queue" = queue.Queue(maxsize=bound)
        self.closed = threading.Event()


class Sender:
    def __init__(self, chan: _Chan):
        self._c = chan

    def send(self, item) -> None:
        if self._c.closed.is_set():
            raise Disconnected()
        self._c.q.put(item)

    def drop(self) -> None:
        self._c.closed.set()


class Receiver:
    def __init__(self, chan: _Chan):
        self._c = chan

    def recv(self, timeout: Optional[float] = None):
        """Blocks like crossbeam's recv(); raises Disconnected once empty and closed."""
        waited = 0.0
        while True:
            try:
                return self._c.q.get(timeout=0.005)
            except queue.Empty:
                if self._c.closed.is_set() and self._c.q.empty():
                    raise Disconnected()
                waited += 0.005
                if timeout is not None and waited >= timeout:
                    raise queue.Empty()

    def try_recv(self):
        try:
            return self._c.q.get_nowait()
        except queue.Empty:
            if self._c.closed.is_set():
                raise Disconnected()
            raise


def unbounded() -> Tuple[Sender, Receiver]:
    c = _Chan(0)
    return Sender(c), Receiver(c)


def bounded(n: int) -> Tuple[Sender, Receiver]:
    c = _Chan(max(1, n))
    return Sender(c), Receiver(c)


class Stretcher:
    """concurrent vocoder for one channel of audio (stretcher.rs:11).

    Stretcher::new(spec, input, factor, amplitude, pitch_multiple, window, buffer_dur,
    frequency_kernel_src) -- `window` is given by its length (the engine builds the Hann
    window of main.rs:131 itself); `frequency_kernel_src` is the path of a kernel file (.cu / .c / .cpp
    source, or a .so exporting the C-ABI `apply`, see include/rocoder_b200.h) that is built, watched
    and hot-swapped as ReFFT::new does (fft.rs:31-35, hotswapper.rs:14-33), or a Python callable.
    """

    def __init__(self, spec: AudioSpec, input: Receiver, factor: float, amplitude: float, pitch_multiple: int,
                 window_len, buffer_dur: float = 1.0, frequency_kernel_src=None, *, channel: int = 0,
                 seed: int = 0, device: int = 0, flags: int = 0):
        assert pitch_multiple != 0  # stretcher.rs:40
        window = None
        if not isinstance(window_len, (int, np.integer)):   # `window: Vec<f32>` given explicitly
            window = np.ascontiguousarray(window_len, dtype=np.float32)
            window_len = int(window.size)
        self.spec = spec
        self.input = input
        self.window_len = window_len
        self.buffer_dur = buffer_dur
        self.channel = channel
        self._engine = Engine(1, spec.sample_rate, window_len, factor, amplitude, pitch_multiple, buffer_dur, seed,
                              device, flags, first_channel=channel)
        if window is not None:
            self._engine.set_window(window)
        p = self._engine.plan
        self.half_window_len = p["half_window_len"]
        self.samples_needed_per_window = p["samples_needed_per_window"]
        self.sample_step_len = p["sample_step_len"]
        self.corrected_amp_factor = p["corrected_amp_factor"]
        self.frames_per_call = p["frames_per_call"]
        self.done = False
        self._sent = 0       # samples handed to the engine
        self._consumed = 0   # samples dropped by truncate_front (stretcher.rs:105-106)
        self._closed = False
        if frequency_kernel_src is not None:
            if callable(frequency_kernel_src):
                self._engine.push_frequency_kernel(frequency_kernel_src)
            else:
                self._engine.watch_frequency_kernel(str(frequency_kernel_src))

    # stretcher.rs:78
    def is_done(self) -> bool:
        return self.done

    # stretcher.rs:82-85
    def channel_bound(self) -> int:
        return channel_bound(self.window_len, self.spec.sample_rate, self.buffer_dur)

    def input_buf_len(self) -> int:
        return self._sent - self._consumed

    # stretcher.rs:123-135
    def ensure_input_samples_available(self, n: int) -> None:
        while self.input_buf_len() < n:
            try:
                chunk = np.asarray(self.input.recv(), dtype=np.float32)
                self._engine.send(0, chunk)
                self._sent += chunk.size
            except Disconnected:
                if not self._closed:
                    self._engine.close_input(0)
                    self._closed = True
                self._sent = self._consumed + n  # input_buf.resize(n, 0.0)
                self.done = True

    # stretcher.rs:87-121
    def next_window(self) -> np.ndarray:
        for _ in range(self.frames_per_call):
            self.ensure_input_samples_available(self.window_len)
            self._consumed += self.sample_step_len
        return self._engine.next_window(0)

    def close(self) -> None:
        self._engine.close()


class AudioBus:
    """audio.rs:142-172"""

    def __init__(self, spec: AudioSpec, channels: List[Receiver], expected_total_samples: Optional[int]):
        self.spec = spec
        self.channels = channels
        self.expected_total_samples = expected_total_samples

    def into_audio(self) -> List[np.ndarray]:
        data: List[List[np.ndarray]] = [[] for _ in self.channels]
        live = [True] * len(self.channels)
        while any(live):
            for i, rx in enumerate(self.channels):
                if not live[i]:
                    continue
                try:
                    data[i].append(rx.recv(timeout=0.005))
                except queue.Empty:
                    pass
                except Disconnected:
                    live[i] = False
        return [np.concatenate(d) if d else np.zeros(0, np.float32) for d in data]


class StretcherProcessor:
    """stretcher_processor.rs:22-75: one thread drives every channel's Stretcher."""

    def __init__(self, channels: List[Tuple[Sender, Stretcher]]):
        self.channels = channels
        self._shutdown = threading.Event()
        self.finished = threading.Event()
        self._thread: Optional[threading.Thread] = None

    @staticmethod
    def new(channel_stretchers: List[Stretcher], expected_total_samples: Optional[int]):
        spec = channel_stretchers[0].spec
        chans, receivers = [], []
        for st in channel_stretchers:
            tx, rx = bounded(st.channel_bound())
            chans.append((tx, st))
            receivers.append(rx)
        return StretcherProcessor(chans), AudioBus(spec, receivers, expected_total_samples)

    def start(self):
        def body():
            try:
                while not self._shutdown.is_set():
                    stop = False
                    for output, stretcher in self.channels:
                        if stretcher.is_done():
                            stop = True  # "assuming each stretcher finishes at the same time"
                            break
                        output.send(stretcher.next_window())
                    if stop:
                        break
            finally:
                for output, _ in self.channels:
                    output.drop()
                self.finished.set()

        self._thread = threading.Thread(target=body, daemon=True)
        self._thread.start()
        return self

    def shutdown(self):
        self._shutdown.set()

    def join(self):
        if self._thread:
            self._thread.join()


def stretch(audio: np.ndarray, sample_rate: int, factor: float = 1.0, pitch_multiple: int = 1,
            window_len: int = 16384, amplitude: float = 1.0, seed: int = 0, device: int = 0,
            flags: int = 0) -> np.ndarray:
    """File-mode run of main.rs:124-160 for planar audio [channels, samples] (batched path)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    if audio.ndim == 1:
        audio = audio[None, :]
    with Engine(audio.shape[0], sample_rate, window_len, factor, amplitude, pitch_multiple, seed=seed,
                device=device, flags=flags) as e:
        return e.process_all(audio)
