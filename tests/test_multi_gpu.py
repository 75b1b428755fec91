"""Time sharding across real GPUs over NCCL (runs whenever the box has at least two GPUs): 2 ranks stretch one
long input with the carried half window exchanged once; the concatenated result equals the
single-GPU run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _log(tmp, rank, msg):
    with open(os.path.join(tmp, f"rank{rank}.log"), "a") as fh:
        fh.write(msg + "\n")
    dbg = os.environ.get("RCG_MGPU_LOG")
    if dbg:
        with open(dbg, "a") as fh:
            fh.write(f"[rank {rank}] {msg}\n")


def _worker(rank, world, port, f, p, W, n, tmp):
    import datetime
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ol.ROOT)
    import rocoder_b200 as rb
    from rocoder_b200.sharding import TimeShardedStretch, input_span
    torch.cuda.set_device(rank)
    _log(tmp, rank, f"start f={f} p={p} W={W} n={n}")
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=90))
    _log(tmp, rank, "process group up")
    try:
        rng = np.random.default_rng(9)
        x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
        eng = rb.Engine(2, 48000, W, f, 1.0, p, seed=5, device=rank)
        calls = eng.expected_counts(n)[0]
        job = TimeShardedStretch(eng, eng.plan, 2, calls, rank, world, torch.device("cuda", rank))
        s0, s1 = input_span(eng.plan, job.c0, job.c1, n)
        seg = torch.from_numpy(np.ascontiguousarray(x[:, s0:s1])).cuda()
        out = torch.full((2, max(job.out_len(), 2)), float("nan"), device="cuda")
        _log(tmp, rank, f"calls [{job.c0},{job.c1}) span [{s0},{s1}) out {job.out_len()}")
        job.run(seg, s0, out, stream=torch.cuda.current_stream().cuda_stream)
        _log(tmp, rank, "run issued")
        torch.cuda.synchronize()
        _log(tmp, rank, "synchronized")
        np.save(os.path.join(tmp, f"out{rank}.npy"), out.cpu().numpy()[:, :job.out_len()])
        dist.barrier()
        _log(tmp, rank, "done")
        dist.destroy_process_group()
    except BaseException as exc:  # never hang the peer in a collective: die loudly instead
        _log(tmp, rank, f"FAILED: {exc!r}")
        os._exit(1)


@pytest.mark.timeout(150)
@pytest.mark.parametrize("f,p,W,n", [(8.0, 2, 16384, 400000), (3.0, -3, 4096, 200000), (2.0, 1, 1024, 90000)])
def test_two_gpus_equal_one(tmp_path, f, p, W, n):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import rocoder_b200 as rb
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), f, p, W, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(9)
    x = ((rng.random((2, n), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)
    with rb.Engine(2, 48000, W, f, 1.0, p, seed=5) as e:
        whole = e.process_all(x)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got, whole)
