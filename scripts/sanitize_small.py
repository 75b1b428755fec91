#!/usr/bin/env python
"""Small runs of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck):
fused full-rate, half-rate, negative pitch, W=65536 CTA pair, plugin paths, boundary, fade."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rocoder_b200 as rb

rng = np.random.default_rng(0)
def inp(ch, n): return ((rng.random((ch, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
cases = [(2, 40000, 2.0, 1, 16384), (2, 40000, 4.0, 2, 16384), (2, 700000, 2.0, 2, 16384), (2, 650000, 2.0, -2, 16384), (2, 30000, 3.0, -3, 4096), (2, 9000, 2.0, 4, 1024),
         (1, 3000, 2.0, 2, 64), (2, 200000, 2.0, 1, 65536), (2, 200000, 2.0, -2, 65536)]
for ch, n, f, p, W in cases:
    with rb.Engine(ch, 48000, W, f, 1.0, p, seed=1) as e:
        out = e.process_all(inp(ch, n))
        assert np.isfinite(out).all()
        # deferred boundary + streaming
        x = inp(ch, n)
        dx = torch.from_numpy(x).cuda()
        calls, _, out_len = e.expected_counts(n)
        if calls >= 2:
            H = W // 2
            o = torch.zeros((ch, out_len + (out_len & 1)), device="cuda")
            head, t0, t1 = (torch.zeros(ch, H, device="cuda") for _ in range(3))
            olen = e.plan["out_len_per_call"]
            c1 = calls // 2
            e.process_segment(dx, dx.stride(0), 0, n, o, o.stride(0), 0, c1, tail_out=t0, stream=1)
            e.process_segment(dx, dx.stride(0), 0, n, o[:, c1 * olen:], o.stride(0), c1, calls - c1, tail_out=t1, head_out=head,
                              defer_first=True, stream=1)
            e.finish_boundary(o[:, c1 * olen:], o.stride(0), c1, head, t0, stream=1)
            torch.cuda.synchronize()
    print("ok", ch, n, f, p, W, flush=True)
with rb.Engine(2, 48000, 1024, 2.0, 1.0, 1, seed=1) as e:
    e.push_frequency_kernel(lambda ms, s: s * np.float32(2.0))
    e.process_all(inp(2, 9000))
with rb.Engine(2, 48000, 1024, 2.0, 1.0, 2, seed=1) as e:
    e.push_device_frequency_kernel("__device__ float2 apply(size_t t, unsigned b, unsigned n, const float2* s) { return s[b]; }")
    e.process_all(inp(2, 9000))
for W, n in ((16384, 120000), (65536, 300000)):   # frequency-kernel path at the large windows (three lanes per CTA; CTA pair)
    with rb.Engine(2, 48000, W, 2.0, 1.0, 1, seed=1) as e:
        e.push_device_frequency_kernel("__device__ float2 apply(size_t t, unsigned b, unsigned n, const float2* s) { return s[b]; }")
        assert np.isfinite(e.process_all(inp(2, n))).all()
    print("ok plugin", W, flush=True)
LOCAL = "__device__ float2 apply(size_t t, unsigned b, unsigned n, float2 x) { return make_float2(x.x * 0.5f, x.y * (b & 1 ? 2.0f : 1.0f)); }"
for W, n, p in ((16384, 120000, 2), (16384, 120000, -2), (1024, 9000, 1), (64, 3000, 3)):   # bin-local kernels fused by NVRTC
    with rb.Engine(2, 48000, W, 2.0, 1.0, p, seed=1) as e:
        e.push_local_device_frequency_kernel(LOCAL)
        assert np.isfinite(e.process_all(inp(2, n))).all()
    print("ok fused local plugin", W, p, flush=True)
d = torch.randn(2, 50000, device="cuda")
rb.fade_in_out(d, 0.25, 48000, 50000, stream=1)
torch.cuda.synchronize()
print("sanitize runs complete")
