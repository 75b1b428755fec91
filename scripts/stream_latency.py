#!/usr/bin/env python
"""Latency of the streaming path: Stretcher::next_window one call per launch (max_lookahead_calls=1)
versus look-ahead batching.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rocoder_b200 as rb
from rocoder_b200.signals import noise_chirp

res = {}
for look in (1, 8, 64):
    x = noise_chirp(48000 * 60, 48000, 1, seed=0)[0]
    with rb.Engine(1, 48000, 16384, 8.0, 1.0, 2, seed=1, max_lookahead_calls=look) as e:
        e.send(0, x); e.close_input(0)
        for _ in range(70): e.next_window(0)   # warm up
        t0 = time.perf_counter(); n = 0
        while not e.is_done(0) and n < 640:
            e.next_window(0); n += 1
        dt = time.perf_counter() - t0
    res[f"lookahead_{look}"] = {"windows": n, "us_per_window": dt / n * 1e6, "window_audio_ms": 16384 / 48000 * 1e3,
                                "x_realtime": (16384 / 48000) / (dt / n)}
print(json.dumps({"streaming_next_window_W16384_f8_p2": res}))
