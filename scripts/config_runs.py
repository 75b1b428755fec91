#!/usr/bin/env python
"""Device-resident timing of the five BASELINE.json configs on 1..N GPUs.

  python scripts/config_runs.py                       # 1 GPU, all configs
  torchrun --nproc-per-node N scripts/config_runs.py  # configs 4 (time-sharded) and 5 (channel-sharded)

Inputs are synthesised on the device (white noise + chirp); one JSON line per config from rank 0.
Kernel time is CUDA-event time around the engine calls, max over ranks."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocoder_b200 as rb
from rocoder_b200.sharding import TimeShardedStretch, input_span

DEVICE_X2 = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    float2 v = spectrum[bin];
    return make_float2(v.x * 2.0f, v.y * 2.0f);
}
"""


LOCAL_X2 = """
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {
    return make_float2(x.x * 2.0f, x.y * 2.0f);
}
"""


def synth(channels, n, rate, dev, seed, offset=0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = (torch.arange(n, device=dev, dtype=torch.float64) + offset) / rate
    T = 600.0
    ph = 2 * math.pi * (20.0 * t + (20000.0 - 20.0) * t * t / (2 * T))
    out = torch.empty((channels, n), device=dev, dtype=torch.float32)
    for c in range(channels):
        out[c] = 0.25 * (torch.rand(n, device=dev, generator=g) * 2 - 1) + 0.25 * torch.sin(ph + c * math.pi / 4).float()
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.current_stream().cuda_stream
    configs = [
        dict(name="config1", secs=10, rate=44100, ch=2, f=1.0, p=1, W=16384),
        dict(name="config2", secs=600, rate=48000, ch=2, f=8.0, p=2, W=16384),
        dict(name="config3_w1024", secs=600, rate=48000, ch=2, f=4.0, p=1, W=1024),
        dict(name="config3_w4096", secs=600, rate=48000, ch=2, f=4.0, p=1, W=4096),
        dict(name="config3_w16384", secs=600, rate=48000, ch=2, f=4.0, p=1, W=16384),
        dict(name="config3_w65536", secs=600, rate=48000, ch=2, f=4.0, p=1, W=65536),
        dict(name="config4", secs=3600, rate=48000, ch=2, f=20.0, p=-3, W=16384, shard="time"),
        dict(name="config5", secs=60, rate=96000, ch=8, f=2.0, p=1, W=16384, shard="channel", plugin=True),
    ]
    only = os.environ.get("RCG_CONFIGS")
    for cfg in configs:
        if only and cfg["name"] not in only.split(","):
            continue
        shard = cfg.get("shard")
        if world > 1 and not shard:
            continue
        n = cfg["secs"] * cfg["rate"]
        ch = cfg["ch"]
        first_channel = 0
        if shard == "channel" and world > 1:
            per = (ch + world - 1) // world
            first_channel = rank * per
            ch = max(0, min(per, cfg["ch"] - first_channel))
        eng = rb.Engine(max(ch, 1), cfg["rate"], cfg["W"], cfg["f"], 1.0, cfg["p"], seed=1234, device=local,
                        first_channel=first_channel)
        if cfg.get("plugin"):
            if os.environ.get("RCG_PLUGIN_FORM", "local") == "local":   # bin-local form: fused into the stretch kernel
                eng.push_local_device_frequency_kernel(LOCAL_X2)
            else:
                eng.push_device_frequency_kernel(DEVICE_X2)
        calls, frames, out_len = eng.expected_counts(n)
        if shard == "time" and world > 1:
            job = TimeShardedStretch(eng, eng.plan, ch, calls, rank, world, dev)
            s0, s1 = input_span(eng.plan, job.c0, job.c1, n)
            x = synth(ch, s1 - s0, cfg["rate"], dev, 1, offset=s0)
            out = torch.empty((ch, job.out_len()), device=dev, dtype=torch.float32)
            run = lambda: job.run(x, s0, out, stream=st)
            my_out = job.out_len() * ch
        else:
            x = synth(max(ch, 1), n, cfg["rate"], dev, 1 + first_channel)
            out = torch.empty((max(ch, 1), out_len), device=dev, dtype=torch.float32)
            run = lambda: eng.process_segment(x, x.stride(0), 0, n, out, out.stride(0), 0, calls, stream=st)
            my_out = out_len * ch
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot = torch.tensor([ms, float(my_out)], device=dev, dtype=torch.float64)
        if dist is not None:
            mx = tot.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            ms, total_out = float(mx[0]), float(tot[1])
        else:
            total_out = float(my_out)
        bytes_alg = 4.0 * (cfg["ch"] * n + total_out)
        # end to end through rcg_process_all with pinned host buffers (1 GPU, outputs up to 4 GB)
        e2e_ms = None
        if os.environ.get("RCG_E2E") and world == 1 and total_out * 4 <= 4e9:
            import time
            hin, hout = rb.PinnedBuffer((ch, n)), rb.PinnedBuffer((ch, out_len))
            hin.array[:] = x.cpu().numpy()
            for _ in range(2):
                eng.process_all_into(hin.array, hout.array)
            t0 = time.perf_counter()
            for _ in range(3):
                eng.process_all_into(hin.array, hout.array)
            e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
            del hin, hout
        if rank == 0:
            print(json.dumps({"config": cfg["name"], "n_gpus": world, "W": cfg["W"], "f": cfg["f"], "p": cfg["p"],
                              "channels": cfg["ch"], "seconds_of_audio": cfg["secs"], "frames": frames * cfg["ch"],
                              "out_samples": total_out, "ms": ms, "out_Msamples_s": total_out / ms / 1e3,
                              "realtime_x": cfg["secs"] / (ms * 1e-3), "hbm_GBs_algorithmic": bytes_alg / ms / 1e6,
                              "sharding": shard if world > 1 else "none", "e2e_ms": e2e_ms, "finite": bool(torch.isfinite(out[:, ::997]).all())}),
                  flush=True)
        del x, out, eng
        torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
