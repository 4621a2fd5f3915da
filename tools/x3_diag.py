"""Diagnostics of the bf16x3 mode (GPU): where the error of the three-plane product comes from, and per-parameter gradient
errors of a train step against the CPU oracle in "f32" and "bf16x3" mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "music-fader-nets_b200"), os.path.join(ROOT, "tests")]
import torch
import fadernets_b200 as fn
from fadernets_b200 import ops_x3 as ox, trainer_gmm
from oracle import fader_oracle as fo

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for K in (256, 1024, 16384):
    M = N = 256
    A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    As = ox.split_bf16(A, M, K, K, 1); Bs = ox.split_bf16(B, N, K, K, 1)
    C = torch.empty(M, N, device=dev)
    ox.tc_gemm_x3(As[0], 0, As[1], As[2], 0, Bs[0], 0, Bs[1], Bs[2], 0, C, 0, N, None, M, N, K)
    ref = A.double() @ B.double().t()
    Av = (As[0][:, :K].double() + As[0][:, As[2]:As[2] + K].double()); Bv = (Bs[0][:, :K].double() + Bs[0][:, Bs[2]:Bs[2] + K].double())
    ah, al = As[0][:, :K].double(), As[0][:, As[2]:As[2] + K].double()
    bh, bl = Bs[0][:, :K].double(), Bs[0][:, Bs[2]:Bs[2] + K].double()
    ref3 = ah @ bh.t() + al @ bh.t() + ah @ bl.t()
    rms = ref.pow(2).mean().sqrt()
    f32 = (A @ B.t()).double()
    print(f"K={K}: rms(result)={rms:.2f}  x3-vs-fp64 {((C.double()-ref).pow(2).mean().sqrt()/rms):.2e}  "
          f"x3-vs-exact-3-plane {((C.double()-ref3).pow(2).mean().sqrt()/rms):.2e}  planes-vs-fp64 {((ref3-ref).pow(2).mean().sqrt()/rms):.2e}  "
          f"torch-fp32 {((f32-ref).pow(2).mean().sqrt()/rms):.2e}  max x3 {((C.double()-ref).abs().max()/rms):.2e}")

variant, H, Z, K, B, T = "gmvae", 256, 128, 2, 16, 24
w = fo.init_weights(H, Z, variant, K, seed=5)
d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=6, pad_tail=True)
gg = torch.Generator().manual_seed(8)
er, en = torch.randn(B, Z, generator=gg), torch.randn(B, Z, generator=gg)
scal, grads, res = fo.loss_and_grads(w, variant, (d, r, n, c, rd, nd), er, en, 20000, 0.2)
for prec in ("f32", "bf16x3", "bf16"):
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    m.load_state_dict(w); m = m.to(dev).train().set_precision(prec)
    it = iter((er, en)); m._draw_eps = lambda B_, Z_, d_: next(it).to(d_); m.host_rng = False
    opt = fn.FusedAdam(m, lr=1e-3); opt.zero_grad()
    trainer_gmm.configure(m, opt, {"beta": 0.2})
    loss, terms, l_r, l_n = trainer_gmm._forward_losses(20000, d.to(dev), r.to(dev), n.to(dev), d.to(dev), r.to(dev), n.to(dev), c.to(dev), rd, nd, False, None)
    loss.backward()
    errs = []
    for k, ref in grads.items():
        got = dict(m.named_parameters())[k].grad.cpu()
        errs.append((float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-6), k))
    errs.sort(reverse=True)
    print(prec, "loss err", abs(float(loss) - float(scal["loss"])) / abs(float(scal["loss"])), [(k, f"{e:.2e}") for e, k in errs[:6]])
