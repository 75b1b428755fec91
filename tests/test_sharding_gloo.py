"""world_size-2 (and 3) gloo runs of the time-sharding host logic on CPU: partition, the one
tail exchange, the boundary fix-up.  The segment backend here is the host emulation of the kernel
phases (test infrastructure); on GPUs the same TimeShardedStretch drives the C ABI over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from rocoder_b200.sharding import TimeShardedStretch, input_span, shard_calls


class EmulatorBackend:
    def __init__(self, f, p, W, seed):
        self.f, self.p, self.W, self.seed = f, p, W, seed

    def process_segment(self, d_in, in_stride, in_origin, in_valid, d_out, out_stride, first_call, n_calls,
                        tail_in=None, tail_out=None, head_out=None, defer_first=False, stream=None):
        x = d_in.numpy()
        ol.emul_segment(x, in_origin, d_out.numpy(), self.f, self.p, self.W, first_call, n_calls, seed=self.seed,
                        tail_in=None if tail_in is None else tail_in.numpy(),
                        tail_out=None if tail_out is None else tail_out.numpy(),
                        head_out=None if head_out is None else head_out.numpy(), defer_first=defer_first)

    def finish_boundary(self, d_out, out_stride, first_call, head, tail_prev, stream=None):
        ol.emul_boundary(d_out.numpy(), self.f, self.p, self.W, first_call, head.numpy(), tail_prev.numpy())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, f, p, W, n, tmp):
    sys.path.insert(0, os.path.join(ol.ROOT, "tests"))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(9)
        x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
        calls = ol.expected_counts(n, f, p, W)[0]
        c = ol.constants(f, p, W)
        plan = dict(c)
        job = TimeShardedStretch(EmulatorBackend(f, p, W, 5), plan, 2, calls, rank, world, "cpu")
        s0, s1 = input_span(plan, job.c0, job.c1, n)
        seg = torch.from_numpy(np.ascontiguousarray(x[:, s0:s1]))
        out = torch.full((2, max(job.out_len(), 1)), float("nan"))
        job.run(seg, s0, out)
        np.save(os.path.join(tmp, f"out{rank}.npy"), out.numpy()[:, :job.out_len()])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,f,p,W,n", [(2, 4.0, 2, 512, 30000), (2, 3.0, -3, 1024, 40000), (3, 2.0, 1, 256, 9000),
                                           (2, 1.0, 1, 256, 300)])
def test_time_sharded_equals_single(tmp_path, world, f, p, W, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, f, p, W, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(9)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    whole = ol.emul_process_all(x, f, p, W, seed=5)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)], axis=1)
    assert got.shape == whole.shape
    assert np.array_equal(got, whole)


def test_shard_calls_partition():
    for total in (0, 1, 7, 8, 14055):
        for world in (1, 2, 3, 8):
            r = shard_calls(total, world)
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
