"""Host-side data-parallel logic on CPU: world_size-2 gloo processes.  The CUDA kernels are not
involved (no GPU here); this covers sharding, the gradient all-reduce hook and weight broadcast."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fadernets_b200 import parallel
    import fadernets_b200 as fn
    # gradient all-reduce == mean over ranks
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    hook = parallel.GradAllReduce()
    hook(g)
    ok_mean = torch.allclose(g, torch.arange(10, dtype=torch.float32) * (1 + 2) / 2)
    # weight broadcast: different seeds -> identical weights afterwards
    torch.manual_seed(100 + rank)
    m = fn.MusicAttrRegVAE(342, 3, 16, 24, 8, 4, 32)
    parallel.broadcast_parameters(m, src=0)
    s = float(sum(p.double().sum() for p in m.state_dict().values()))
    sums = [None] * world
    dist.all_gather_object(sums, s)
    # overlapped, bucketed all-reduce: bucket 0 (decoder-side parameters) goes out from the post-accumulate hooks while
    # backward is still running, bucket 1 (the encoders) in step(); SUM collectives + 1/world folded into the loss
    flat, grad = m.flatten_parameters_()
    ov = parallel.OverlappedGradAllReduce(m)
    names = [n for n, _ in m.live_parameters()]
    first_enc = min(i for i, n in enumerate(names) if n.startswith(parallel.ENCODER_PREFIXES))
    ok_order = all(n.startswith(parallel.ENCODER_PREFIXES) for n in names[first_enc:]) and 0 < ov.split < grad.numel()
    grad.zero_()
    loss = sum(((rank + 1) * (i + 1)) * p.sum() for i, (_, p) in enumerate(m.live_parameters()))
    (loss * ov.loss_scale).backward()
    issued_early = ov._handle is not None                      # bucket 0 was issued from the hooks, before the sync call
    ov(grad)
    expect = torch.cat([torch.full((((p.numel() + 3) // 4) * 4,), (i + 1) * (1 + 2) / 2.0) * torch.cat(
        [torch.ones(p.numel()), torch.zeros(((p.numel() + 3) // 4) * 4 - p.numel())]) for i, (_, p) in enumerate(m.live_parameters())])
    ok_overlap = ok_order and issued_early and torch.allclose(grad, expect) and ov.calls == 1
    # second step re-arms the hooks
    grad.zero_()
    (loss * ov.loss_scale).backward() if False else (sum(p.sum() for _, p in m.live_parameters()) * ov.loss_scale).backward()
    ov(grad)
    ok_overlap = ok_overlap and ov.calls == 2 and bool((grad[:5] == 1.0).all())
    # shards tile the batch
    lo, hi = parallel.shard_bounds(11, rank, world)
    q.put((rank, ok_mean and ok_overlap, sums, (lo, hi), hook.calls))
    dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2][0] == res[0][2][1], "weights differ after broadcast"
    assert res[0][3] == (0, 6) and res[1][3] == (6, 11)
    assert all(r[4] == 1 for r in res)


def test_shard_batch_layout():
    sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200"))
    from fadernets_b200 import parallel
    batch = (torch.arange(8).view(8, 1), torch.zeros(8, 2), torch.ones(8, 3), torch.rand(8, 24), np.arange(8.0), np.arange(8.0))
    parts = [parallel.shard_batch(batch, r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p[0] for p in parts]), batch[0])
    assert np.array_equal(np.concatenate([p[4] for p in parts]), batch[4])
