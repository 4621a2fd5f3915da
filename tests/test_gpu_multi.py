"""Two-GPU (NCCL) parity of the data-parallel step against ONE GPU on the full batch: the overlapped, bucketed gradient
all-reduce (1/world folded into backward) and the exact global-batch latent regulariser (reference
trainer_gmm.py:199-217 evaluated at batch B * world).  Needs two visible GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(dev, seed=0):
    import fadernets_b200 as fn
    torch.manual_seed(seed)
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, 64, 16, 32, n_component=2)
    return m.to(dev).train()


def _batch(B, T):
    g = torch.Generator().manual_seed(3)
    d = torch.randint(2, 342, (B, T), generator=g); r = torch.randint(0, 3, (B, T), generator=g)
    n = torch.randint(0, 16, (B, T), generator=g); c = torch.rand(B, 24, generator=g)
    eps = torch.randn(2, B, 16, generator=g)
    return d, r, n, c, (r == 1).double().mean(1), n.double().mean(1), eps


def _losses_and_grads(m, opt_sync, d, r, n, c, rd, nd, eps, dev):
    from fadernets_b200 import trainer_gmm, _steps
    it = iter((eps[0], eps[1]))
    m._draw_eps = lambda B_, Z_, d_: next(it).to(d_)
    m.host_rng = False
    trainer_gmm.configure(m, None, {"beta": 0.2})
    m.zero_grad_flat()
    loss, terms, l_r, l_n = trainer_gmm._forward_losses(20000, d.to(dev), r.to(dev), n.to(dev), d.to(dev), r.to(dev), n.to(dev),
                                                        c.to(dev), rd, nd, False, None)
    scale = opt_sync.loss_scale if opt_sync is not None else 1.0
    (loss * scale).backward()
    flat, grad = m.flatten_parameters_()
    if opt_sync is not None:
        opt_sync(grad)
    torch.cuda.synchronize()
    return float(loss), float(l_r), float(l_n), grad.detach().cpu().clone()


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "music-fader-nets_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fadernets_b200 import parallel
    B, T = 8, 12
    d, r, n, c, rd, nd, eps = _batch(B, T)
    m = _build(dev)
    parallel.broadcast_parameters(m, src=0)
    lo, hi = parallel.shard_bounds(B, rank, world)
    ov = parallel.OverlappedGradAllReduce(m)
    parallel.enable_global_latent_reg(True)
    loss, l_r, l_n, grad = _losses_and_grads(m, ov, d[lo:hi], r[lo:hi], n[lo:hi], c[lo:hi], rd[lo:hi], nd[lo:hi], eps[:, lo:hi], dev)
    early = ov.calls == 1
    out = dict(rank=rank, l_r=l_r, l_n=l_n, grad=grad, early=early)
    if rank == 0:
        parallel.enable_global_latent_reg(False)
        m1 = _build(dev)
        loss1, l_r1, l_n1, grad1 = _losses_and_grads(m1, None, d, r, n, c, rd, nd, eps, dev)
        out.update(l_r1=l_r1, l_n1=l_n1, grad1=grad1)
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_step_equals_single_gpu_full_batch(lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert torch.equal(r0["grad"], r1["grad"]), "ranks hold different averaged gradients"
    # the global-batch latent regulariser has the same value on every rank, equal to the single-GPU full-batch value
    assert abs(r0["l_r"] - r1["l_r"]) < 1e-6 and abs(r0["l_r"] - r0["l_r1"]) < 1e-5 * max(1.0, abs(r0["l_r1"]))
    assert abs(r0["l_n"] - r0["l_n1"]) < 1e-5 * max(1.0, abs(r0["l_n1"]))
    # averaged data-parallel gradient == gradient of the full batch on one GPU (fp32 path: re-association only)
    g, g1 = r0["grad"].double(), r0["grad1"].double()
    err = float((g - g1).abs().max()) / float(g1.abs().max())
    assert err < 2e-4, err
