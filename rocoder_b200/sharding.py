"""Time sharding of one long stretch job across the GPUs of one box.

Frame j reads in[j*step : j*step+W) and its own counter-based phases; output block m needs
frames m and m-1 only (there is no phase carry in the reference, fft.rs:63-68).  So the calls
of a channel split into contiguous ranges, one per rank, with ONE exchange: rank r's last
carried half window (channels x W/2 floats) goes to rank r+1, which adds it into its first
block (rcg_finish_boundary).  One process per GPU; the exchange is a torch.distributed
point-to-point (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Protocol, Tuple


def shard_calls(total_calls: int, world: int) -> List[Tuple[int, int]]:
    """contiguous, near-equal call ranges [c0, c1) per rank (possibly empty at the end)"""
    base, rem = divmod(total_calls, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((c, c + n))
        c += n
    return out


def input_span(plan: dict, c0: int, c1: int, n_in: int) -> Tuple[int, int]:
    """samples [s0, s1) of the channel that calls [c0, c1) read (no halo frame: the carried half
    window arrives from the previous rank)"""
    k, step, W = plan["frames_per_call"], plan["sample_step_len"], plan["window_len"]
    if c1 <= c0:
        return 0, 0
    s0 = min(n_in, c0 * k * step)
    s1 = min(n_in, (c1 * k - 1) * step + W)
    return s0, max(s0, s1)


class SegmentBackend(Protocol):
    def process_segment(self, d_in, in_stride, in_origin, in_valid, d_out, out_stride, first_call, n_calls,
                        tail_in=None, tail_out=None, head_out=None, defer_first=False, stream=None) -> None: ...

    def finish_boundary(self, d_out, out_stride, first_call, head, tail_prev, stream=None) -> None: ...


class TimeShardedStretch:
    """One rank's share of a time-sharded job.  Buffers are torch tensors on the rank's device."""

    def __init__(self, backend: SegmentBackend, plan: dict, channels: int, total_calls: int, rank: int, world: int,
                 device, process_group=None):
        import torch
        self.torch = torch
        self.backend, self.plan, self.channels = backend, plan, channels
        self.rank, self.world, self.group = rank, world, process_group
        self.c0, self.c1 = shard_calls(total_calls, world)[rank]
        H = plan["half_window_len"]
        self.head = torch.zeros(channels, H, dtype=torch.float32, device=device)
        self.tail = torch.zeros(channels, H, dtype=torch.float32, device=device)
        self.prev_tail = torch.zeros(channels, H, dtype=torch.float32, device=device)
        # ranks with no calls neither send nor receive; find the neighbours that have work
        ranges = shard_calls(total_calls, world)
        self.has_work = self.c1 > self.c0
        self.prev_rank = next((r for r in range(rank - 1, -1, -1) if ranges[r][1] > ranges[r][0]), None)
        self.next_rank = next((r for r in range(rank + 1, world) if ranges[r][1] > ranges[r][0]), None)

    def out_len(self) -> int:
        return (self.c1 - self.c0) * self.plan["out_len_per_call"]

    def run(self, seg_in, in_origin: int, out, stream: Optional[int] = None) -> None:
        """seg_in [channels, n] holds the channel samples from in_origin on; out [channels, out_len()].
        `stream` is a cudaStream_t handle; None = torch's current stream, the one torch.distributed orders its
        point-to-point operations against (the engine's private stream would let the send start before the
        stretch kernel has written the carried half window)."""
        if not self.has_work:
            return
        dist = self.torch.distributed
        if stream is None and self.head.is_cuda:
            stream = self.torch.cuda.current_stream(self.head.device).cuda_stream
        defer = self.prev_rank is not None
        self.backend.process_segment(seg_in, seg_in.stride(0), in_origin, seg_in.shape[1], out, out.stride(0),
                                     self.c0, self.c1 - self.c0, tail_out=self.tail, head_out=self.head,
                                     defer_first=defer, stream=stream)
        ops = []
        if self.next_rank is not None:
            ops.append(dist.P2POp(dist.isend, self.tail, self.next_rank, self.group))
        if self.prev_rank is not None:
            ops.append(dist.P2POp(dist.irecv, self.prev_tail, self.prev_rank, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if defer:
            self.backend.finish_boundary(out, out.stride(0), self.c0, self.head, self.prev_tail, stream=stream)
