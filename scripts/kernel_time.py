#!/usr/bin/env python
"""Device-resident timing of the stretch kernel only (BASELINE config 2 by default); used to compare
build variants.  usage: RCG_LIB=path python scripts/kernel_time.py [--pitch 2] [--factor 8] [--window 16384] [--seconds 600]"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rocoder_b200 as rb
from rocoder_b200.signals import noise_chirp

ap = argparse.ArgumentParser()
ap.add_argument("--pitch", type=int, default=2); ap.add_argument("--factor", type=float, default=8.0)
ap.add_argument("--window", type=int, default=16384); ap.add_argument("--seconds", type=int, default=600)
ap.add_argument("--channels", type=int, default=2); ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--tag", default="")
a = ap.parse_args()
n = 48000 * a.seconds
x = noise_chirp(n, 48000, a.channels, seed=0)
e = rb.Engine(a.channels, 48000, a.window, a.factor, 1.0, a.pitch, seed=1234)
calls, frames, out_len = e.expected_counts(n)
d_in = torch.from_numpy(x).cuda()
d_out = torch.empty((a.channels, out_len), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(a.reps + 3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e.process_segment(d_in, d_in.stride(0), 0, n, d_out, d_out.stride(0), 0, calls, stream=st)
    e1.record(); torch.cuda.synchronize()
    if i >= 3: ts.append(e0.elapsed_time(e1))
ms = statistics.median(ts)
print(json.dumps({"tag": a.tag, "lib": os.environ.get("RCG_LIB", "default"), "W": a.window, "f": a.factor, "p": a.pitch,
                  "frames": frames * a.channels, "kernel_ms": ms, "min_ms": min(ts), "out_Msamples_s": out_len * a.channels / ms / 1e3,
                  "us_per_frame_per_sm": ms * 1e3 * 148 / (frames * a.channels), "checksum": float(d_out[:, ::4099].abs().sum()), "geometry": e.launch_geometry()}))
