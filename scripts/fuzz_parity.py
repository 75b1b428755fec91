#!/usr/bin/env python
"""Randomised parity sweep of the CUDA path against the oracle: random window, factor, pitch multiple, input
length and channel count inside the reference's valid domain (1 <= step <= W, p not in {0, -1}).
Index path bit-exact (bypass mode), PCM within the north-star tolerance of the fp64 arbiter, batch == streaming,
batch == time-sharded segments (halo recompute / tail chaining / deferred boundary) at random cut points,
batch == the same job partitioned over a device list behind one engine (exchange and halo variants); one window in
five is an even length that is not a power of two (chirp-z path).
usage: python scripts/fuzz_parity.py [n_cases] [seed]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
import rocoder_b200 as rb

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = []
done = 0
n_stream = n_segments = n_devices = n_anyw = 0
while done < n_cases:
    W = int(2 ** rng.integers(6, 17))
    if rng.random() < 0.2:
        W = 2 * int(rng.integers(32, 700))      # even, usually not a power of two; small enough for the oracle's O(n^2) DFT
        if W & (W - 1) == 0:
            W += 2
        n_anyw += 1
    p = int(rng.choice([1, 1, 2, 2, 3, 4, 5, 6, 8, -2, -2, -3, -3, -4, -5, -7]))
    f = float(np.float32(rng.choice([0.5, 0.75, 1.0, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 8.0, 20.0, rng.uniform(0.5, 12.0)])))
    try:
        c = ol.constants(f, p, W)
    except Exception:
        continue
    step = c["sample_step_len"]
    if not (1 <= step <= W):
        continue
    ch = int(rng.choice([1, 2, 2, 3, 8]))
    frames_target = int(rng.integers(1, 40 if W >= 16384 else (30 if W & (W - 1) else 120)))
    n = int(rng.integers(0, W + frames_target * step + 1))
    if n * c["pitch_shifted_factor"] * ch > 6e6:     # keep the oracle in seconds
        continue
    x = ((rng.random((ch, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    seed = int(rng.integers(0, 2 ** 40))
    tag = dict(W=W, p=p, f=f, ch=ch, n=n, seed=seed)
    try:
        want_b = ol.process_all(x, f, p, W, mode=ol.RO_PHASE_BYPASS)
        ref = ol.process_all(x, f, p, W, seed=seed, precision=64)
        with rb.Engine(ch, 48000, W, f, 1.0, p, flags=rb.RCG_FLAG_BYPASS_RESYNTH) as e:
            got_b = e.process_all(x)
        with rb.Engine(ch, 48000, W, f, 1.0, p, seed=seed) as e:
            got = e.process_all(x)
        ok_b = got_b.shape == want_b.shape and np.array_equal(got_b, want_b)
        err = float(np.abs(got - ref).max()) if ref.size else 0.0
        ok_p = got.shape == ref.shape and np.isfinite(got).all() and err <= 1e-4 and (not ref.any() or ol.snr_db(ref, got) > 100.0)
        ok_s = True
        if done % 4 == 0 and n > 0:      # streaming next_window == batch, bit for bit
            n_stream += 1
            with rb.Engine(ch, 48000, W, f, 1.0, p, seed=seed, max_lookahead_calls=int(rng.integers(1, 6))) as e:
                for cc in range(ch):
                    cut = int(rng.integers(0, n + 1))
                    e.send(cc, x[cc, :cut]); e.send(cc, x[cc, cut:]); e.close_input(cc)
                outs = [[] for _ in range(ch)]
                while not e.is_done(0):
                    for cc in range(ch):
                        outs[cc].append(e.next_window(cc))
                st = np.stack([np.concatenate(o) for o in outs])
            ok_s = st.shape == got.shape and np.array_equal(st, got)
        ok_g = True
        calls = ol.expected_counts(n, f, p, W)[0]
        if done % 3 == 0 and calls >= 3 and got.shape[1] % 2 == 0:
            # time sharding: random cut points; halo recompute, tail chaining and deferred boundary == one shot
            import torch
            n_segments += 1
            k, olen, H = c["frames_per_call"], c["out_len_per_call"], W // 2
            cuts = sorted(set([0, calls] + [int(v) for v in rng.integers(1, calls, size=int(rng.integers(1, 4)))]))
            dx = torch.from_numpy(x).cuda()
            st_ = torch.cuda.current_stream().cuda_stream
            with rb.Engine(ch, 48000, W, f, 1.0, p, seed=seed) as e:
                res = []
                out = torch.full((ch, got.shape[1]), float("nan"), device="cuda")
                for c0, c1 in zip(cuts[:-1], cuts[1:]):
                    s0 = max(c0 * k - 1, 0) * step
                    s0 -= s0 % 2   # float2 loads want an even origin
                    seg = dx[:, s0:].contiguous() if (n - s0) % 2 == 0 else torch.nn.functional.pad(dx[:, s0:], (0, 1)).contiguous()
                    e.process_segment(seg, seg.stride(0), s0, n - s0, out[:, c0 * olen:], out.stride(0), c0, c1 - c0, stream=st_)
                torch.cuda.synchronize(); res.append(out.cpu().numpy())
                out = torch.full((ch, got.shape[1]), float("nan"), device="cuda")
                tails = [torch.zeros(ch, H, device="cuda") for _ in cuts]
                for i, (c0, c1) in enumerate(zip(cuts[:-1], cuts[1:])):
                    e.process_segment(dx, dx.stride(0), 0, n, out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                                      tail_in=tails[i] if c0 else None, tail_out=tails[i + 1], stream=st_)
                torch.cuda.synchronize(); res.append(out.cpu().numpy())
                out = torch.full((ch, got.shape[1]), float("nan"), device="cuda")
                heads = [torch.zeros(ch, H, device="cuda") for _ in cuts]
                tails = [torch.zeros(ch, H, device="cuda") for _ in cuts]
                for i, (c0, c1) in enumerate(zip(cuts[:-1], cuts[1:])):
                    e.process_segment(dx, dx.stride(0), 0, n, out[:, c0 * olen:], out.stride(0), c0, c1 - c0,
                                      tail_out=tails[i], head_out=heads[i], defer_first=c0 > 0, stream=st_)
                for i, c0 in enumerate(cuts[1:-1], start=1):
                    e.finish_boundary(out[:, c0 * olen:], out.stride(0), c0, heads[i], tails[i - 1], stream=st_)
                torch.cuda.synchronize(); res.append(out.cpu().numpy())
            ok_g = all(np.array_equal(r, got) for r in res)
        ok_d = True
        if done % 5 == 0 and n > 0:      # one engine, several devices (the same GPU listed repeatedly): bit-identical
            n_devices += 1
            devs = [0] * int(rng.integers(2, 6))
            fl = rb.RCG_FLAG_SHARD_ALWAYS | (rb.RCG_FLAG_SHARD_HALO if rng.random() < 0.5 else 0)
            with rb.Engine(ch, 48000, W, f, 1.0, p, seed=seed, devices=devs, flags=fl) as e:
                many = e.process_all(x)
            ok_d = many.shape == got.shape and np.array_equal(many, got)
        if not (ok_b and ok_p and ok_s and ok_g and ok_d):
            bad.append(dict(tag, index_exact=bool(ok_b), pcm_ok=bool(ok_p), stream_ok=bool(ok_s), segments_ok=bool(ok_g),
                            devices_ok=bool(ok_d), max_abs=err))
            print("FAIL", json.dumps(bad[-1]), flush=True)
    except Exception as ex:   # noqa
        bad.append(dict(tag, error=repr(ex)[:300]))
        print("ERROR", json.dumps(bad[-1]), flush=True)
    done += 1
    if done % 20 == 0:
        print(f"{done} cases, {len(bad)} failures", flush=True)
print(json.dumps({"cases": done, "streaming_checks": n_stream, "segment_checks": n_segments, "device_list_checks": n_devices,
                  "non_power_of_two_windows": n_anyw, "failures": len(bad), "failed": bad[:20]}))
sys.exit(1 if bad else 0)
