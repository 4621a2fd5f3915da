"""Per-kernel parity (GPU): each C-ABI entry point against a plain torch fp32/fp64 restatement of
the same op on the same seeded inputs.  Tolerances are fp32 re-association only."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def rnd(*shape, seed=0, dev=None, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    lim = atol + rtol * b.abs().max().item() if b.numel() else atol
    assert err <= lim, f"{what}: max abs err {err:.3e} > {lim:.3e}"


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (64, 48, 16), (130, 342, 70), (257, 129, 515), (300, 1536, 64),
                                   (1100, 1200, 40), (33, 16, 0)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_all_layouts(dev, M, N, K, ta, tb):
    from fadernets_b200 import ops
    A = rnd(M, K, seed=1, dev=dev) if not ta else rnd(K, M, seed=1, dev=dev)
    Bm = rnd(K, N, seed=2, dev=dev) if not tb else rnd(N, K, seed=2, dev=dev)
    bias = rnd(N, seed=3, dev=dev)
    C0 = rnd(M, N + 3, seed=4, dev=dev)
    Cm = C0.clone()
    sam, sak = (K, 1) if not ta else (1, M)
    sbk, sbn = (N, 1) if not tb else (1, K)
    ops.gemm(A, 0, sam, sak, Bm, 0, sbk, sbn, Cm, 0, N + 3, bias, M, N, K, accumulate=True)
    Am = A if not ta else A.t()
    Bmm = Bm if not tb else Bm.t()
    ref = C0.clone().double()
    ref[:, :N] += Am.double() @ Bmm.double() + bias.double()
    close(Cm, ref, rtol=2e-5, atol=1e-5, what="gemm")
    assert torch.equal(Cm[:, N:], C0[:, N:]), "wrote outside the tile"


def _torch_gru(gi, h0, w_hh, b_hh, reverse):
    T, B, _ = gi.shape
    H = w_hh.shape[1]
    h = h0
    hs = [None] * T
    for s in range(T):
        t = T - 1 - s if reverse else s
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[t, :, :H] + gh[:, :H])
        z = torch.sigmoid(gi[t, :, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[t, :, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        hs[t] = h
    return torch.stack(hs, 0)


@pytest.mark.parametrize("B,T,H,Vin,Zin", [(3, 5, 16, 7, 4), (70, 9, 64, 342, 12), (64, 6, 256, 16, 8), (5, 1, 32, 3, 8)])
def test_gru_group_forward_backward(dev, B, T, H, Vin, Zin):
    """Three chains in one launch: (emb, reverse, zero h0, final state) / (emb + z-projection, h0) /
    (dense input, h0 = xin[0]) -- forward values and every gradient against torch autograd (fp64)."""
    from fadernets_b200.ops import ChainSpec, GruGroupFn
    k = 1.0 / math.sqrt(H)
    P = {}
    def par(name, *shape, seed):
        P[name] = (rnd(*shape, seed=seed, dev=dev) * k).requires_grad_(True)
        return P[name]
    ids = torch.randint(0, Vin, (T, B), generator=torch.Generator().manual_seed(5)).int().to(dev)
    # chain A: encoder-like
    wa = [par("a_wih", 3 * H, Vin, seed=10), par("a_bih", 3 * H, seed=11), par("a_whh", 3 * H, H, seed=12), par("a_bhh", 3 * H, seed=13)]
    # chain B: sub-decoder-like
    wb = [par("b_wih", 3 * H, Vin + Zin, seed=20), par("b_bih", 3 * H, seed=21), par("b_whh", 3 * H, H, seed=22), par("b_bhh", 3 * H, seed=23)]
    zin = rnd(B, Zin, seed=24, dev=dev).requires_grad_(True)
    h0b = rnd(B, H, seed=25, dev=dev).requires_grad_(True)
    # chain C: second-cell-like
    wc = [par("c_wih", 3 * H, H, seed=30), par("c_bih", 3 * H, seed=31), par("c_whh", 3 * H, H, seed=32), par("c_bhh", 3 * H, seed=33)]
    xin = rnd(T, B, H, seed=34, dev=dev, scale=0.5).requires_grad_(True)
    specs = [ChainSpec(emb_cols=(0, Vin), ids=ids, reverse=True, final=(0, 2)),
             ChainSpec(emb_cols=(0, Vin), ids=ids, z_cols=(Vin, Zin), h0="tensor", want_hs=True),
             ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)]
    fin, hs_b, hs_c = GruGroupFn.apply(specs, B, T, H, (H + 5,), *wa, *wb, zin, h0b, *wc, xin)
    go_f, go_b, go_c = rnd(B, H, seed=40, dev=dev), rnd(T, B, H, seed=41, dev=dev), rnd(T, B, H, seed=42, dev=dev)
    loss = (fin[:, 2:2 + H] * go_f).sum() + (hs_b * go_b).sum() + (hs_c * go_c).sum()
    leaves = wa + wb + [zin, h0b] + wc + [xin]
    grads = torch.autograd.grad(loss, leaves)

    # torch fp64 restatement
    D = [t.detach().double().requires_grad_(True) for t in leaves]
    a_wih, a_bih, a_whh, a_bhh, b_wih, b_bih, b_whh, b_bhh, zin_, h0b_, c_wih, c_bih, c_whh, c_bhh, xin_ = D
    idl = ids.long()
    gi_a = a_wih.t()[idl] + a_bih
    ra = _torch_gru(gi_a, torch.zeros(B, H, dtype=torch.float64, device=dev), a_whh, a_bhh, True)
    gi_b = b_wih[:, :Vin].t()[idl] + (zin_ @ b_wih[:, Vin:].t() + b_bih)[None]
    rb = _torch_gru(gi_b, h0b_, b_whh, b_bhh, False)
    gi_c = xin_ @ c_wih.t() + c_bih
    rc = _torch_gru(gi_c, xin_[0], c_whh, c_bhh, False)
    close(fin[:, 2:2 + H], ra[0], what="final state (reverse chain)")
    close(hs_b, rb, what="hs chain B")
    close(hs_c, rc, what="hs chain C")
    rloss = (ra[0] * go_f.double()).sum() + (rb * go_b.double()).sum() + (rc * go_c.double()).sum()
    rgrads = torch.autograd.grad(rloss, D)
    names = ["a_wih", "a_bih", "a_whh", "a_bhh", "b_wih", "b_bih", "b_whh", "b_bhh", "zin", "h0b", "c_wih", "c_bih",
             "c_whh", "c_bhh", "xin"]
    for nm, g, rg in zip(names, grads, rgrads):
        close(g, rg, rtol=2e-4, atol=1e-5, what="grad " + nm)


def test_softmax_heads_and_nll(dev):
    from fadernets_b200.ops import NllMeanFn, TimeLogSoftmaxFn, VocabLogSoftmaxFn, VocabNllFn
    B, T, V = 5, 7, 342
    x = rnd(T, B, V, seed=1, dev=dev, scale=3).requires_grad_(True)
    tgt = torch.randint(0, V, (B, T), generator=torch.Generator().manual_seed(2)).to(dev)
    out = VocabLogSoftmaxFn.apply(x)
    ref_in = x.detach().double().requires_grad_(True)
    ref = torch.log_softmax(ref_in, -1).permute(1, 0, 2)
    close(out, ref, what="vocab log-softmax")
    go = rnd(B, T, V, seed=3, dev=dev)
    (g,) = torch.autograd.grad((out * go).sum(), x)
    (rg,) = torch.autograd.grad((ref * go.double()).sum(), ref_in, retain_graph=True)
    close(g, rg, what="vocab log-softmax grad")
    # NLL mean through the generic path and through the fused path
    l1 = NllMeanFn.apply(VocabLogSoftmaxFn.apply(x), tgt)
    l2 = VocabNllFn.apply(x, tgt)
    rl = torch.nn.functional.nll_loss(ref.reshape(-1, V), tgt.reshape(-1))
    close(l1, rl, what="nll mean"); close(l2, rl, what="fused nll")
    (g1,) = torch.autograd.grad(l1 * 3.0, x); (g2,) = torch.autograd.grad(l2 * 3.0, x)
    (rg,) = torch.autograd.grad(rl * 3.0, ref_in)
    close(g1, rg, what="nll grad"); close(g2, rg, what="fused nll grad")
    # time-axis soft-max
    for Cc in (3, 16):
        y = rnd(T, B, Cc, seed=4, dev=dev, scale=2).requires_grad_(True)
        o = TimeLogSoftmaxFn.apply(y)
        ry = y.detach().double().requires_grad_(True)
        ro = torch.log_softmax(ry.permute(1, 0, 2), 1)
        close(o, ro, what="time log-softmax")
        gg = rnd(B, T, Cc, seed=5, dev=dev)
        (g,) = torch.autograd.grad((o * gg).sum(), y); (rg,) = torch.autograd.grad((ro * gg.double()).sum(), ry)
        close(g, rg, what="time log-softmax grad")


def test_token_plumbing(dev):
    from fadernets_b200 import ops
    B, T, V = 6, 9, 342
    ids = torch.randint(0, V, (B, T), generator=torch.Generator().manual_seed(1)).to(dev)
    oh = ops.ids_to_onehot(ids, V)
    assert torch.equal(oh, torch.nn.functional.one_hot(ids, V).float())
    assert torch.equal(ops.onehot_to_ids_tm(oh).long(), ids.t())
    sh = ops.ids_to_tm(ids, shift=1, start_token=341)
    assert torch.equal(sh[0].long(), torch.full((B,), 341, device=dev)) and torch.equal(sh[1:].long(), ids.t()[:-1])
    # first-max tie rule of _sampling (gmm_model.py:73-80)
    x = torch.zeros(3, 1, V, device=dev); x[0, 0, 5] = 2; x[0, 0, 300] = 2; x[1, 0, 341] = 1
    assert ops.onehot_to_ids_tm(x).view(-1).tolist() == [5, 341, 0]


@pytest.mark.parametrize("B,Z", [(37, 24), (300, 128), (33, 23), (70, 260), (1500, 128)])
@pytest.mark.parametrize("K", [1, 2, 4])
def test_latent_block(dev, K, B, Z):
    """Latent block against a torch fp64 restatement.  (B, Z) cover the streaming kernels (Z % 4 == 0: one and several
    float4 chunks per lane, more rows than one pass of the grid) and the scalar fallback (Z = 23)."""
    from fadernets_b200.ops import ExpFn, GmKlFn, LatentHeadFn, LatentRegFn, QyXFn, StdKlFn
    from torch.distributions import Normal, kl_divergence
    mu = rnd(B, Z, seed=1, dev=dev).requires_grad_(True)
    pre = rnd(B, Z, seed=2, dev=dev, scale=0.3).requires_grad_(True)
    eps = rnd(B, Z, seed=3, dev=dev)
    mul = rnd(K, Z, seed=4, dev=dev, scale=0.3).requires_grad_(True)
    lvl = torch.full((K, Z), -4.0, device=dev) + rnd(K, Z, seed=5, dev=dev, scale=0.1)
    ylab = torch.randint(0, K, (B,), generator=torch.Generator().manual_seed(6)).to(dev)
    attr = torch.rand(B, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    attr[3] = attr[4]
    for mode in (0, 1):
        scale, z = LatentHeadFn.apply(mu, pre, eps)
        ll, qy, y = QyXFn.apply(z, mul, lvl)
        kl = GmKlFn.apply(mu, scale, mul, lvl, qy, ll, ylab if mode else None, mode)
        sk = StdKlFn.apply(mu, ExpFn.apply(pre))
        lr = LatentRegFn.apply(z, attr.to(dev))
        tot = 1.3 * kl[0] + 0.7 * kl[1] + 0.9 * kl[2] + 0.5 * sk + 2.0 * lr + (ll * 0.01).sum() + (qy * qy).sum()
        g = torch.autograd.grad(tot, [mu, pre, mul])
        # torch restatement (fp64)
        m_, p_, l_ = (t.detach().double().requires_grad_(True) for t in (mu, pre, mul))
        lv_ = lvl.double()
        s_ = p_.exp()
        z_ = m_ + s_ * eps.double()
        d_ = z_[:, None] - l_[None]
        llr = (-0.5 * (d_ * d_ / lv_.exp()[None] + lv_[None] + math.log(2 * math.pi))).sum(-1) + math.log(1.0 / K)
        qyr = torch.softmax(llr, 1)
        q = Normal(m_, s_)
        if mode == 0:
            lat = sum((kl_divergence(q, Normal(l_[k], lv_[k].exp())).mean(-1) * qyr[:, k]).mean() for k in range(K))
            cls = ((qyr * torch.log_softmax(llr, 1)).mean(1) - math.log(1.0 / K)).mean()
            clf = torch.zeros((), dtype=torch.float64, device=dev)
        else:
            lat = kl_divergence(q, Normal(l_[ylab], lv_[ylab].exp())).mean(-1).mean()
            cls = torch.zeros((), dtype=torch.float64, device=dev)
            clf = torch.nn.functional.cross_entropy(qyr, ylab)
        skr = kl_divergence(q, Normal(torch.zeros_like(m_), torch.ones_like(s_))).mean()
        a = attr.numpy()
        sg = torch.sign(torch.from_numpy(np.subtract.outer(a, a)).float()).double().to(dev)
        lrr = ((torch.tanh(z_[:, 0].reshape(-1, 1) - z_[:, 0]) - sg) ** 2).mean()
        close(z, z_, what="z"); close(ll, llr, rtol=1e-5, atol=1e-3, what="logLogit"); close(qy, qyr, rtol=1e-4, atol=4e-4, what="qy")   # logits of magnitude ~Z e^4: one fp32 ulp of the logit is ~5e-4
        assert torch.equal(y, qyr.max(1)[1])
        close(kl[0], lat, what=f"kld_lat mode {mode}"); close(kl[1], cls, what="kld_cls"); close(kl[2], clf, what="clf")
        close(sk, skr, what="std kl"); close(lr, lrr, what="latent reg")
        totr = 1.3 * lat + 0.7 * cls + 0.9 * clf + 0.5 * skr + 2.0 * lrr + (llr * 0.01).sum() + (qyr * qyr).sum()
        rg = torch.autograd.grad(totr, [m_, p_, l_])
        for nm, x, y_ in zip(("mu", "pre", "mu_lookup"), g, rg):
            close(x, y_, rtol=1e-3, atol=1e-5, what=f"grad {nm} mode {mode}")   # the path's fp32 bar is 1e-3


def test_clip_adam_matches_torch(dev):
    import fadernets_b200 as fn
    torch.manual_seed(0)
    m = fn.MusicAttrRegVAE(342, 3, 16, 24, 16, 8, 32).to(dev)
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in m.live_parameters()}
    topt = torch.optim.Adam(list(ref.values()), lr=1e-3)
    opt = fn.FusedAdam(m, lr=1e-3)
    for it in range(3):
        opt.zero_grad()
        for i, (n, p) in enumerate(m.live_parameters()):
            g = rnd(*p.shape, seed=100 * it + i, dev=dev, scale=0.05 * (it + 1))
            p.grad.copy_(g); ref[n].grad = g.clone()
        torch.nn.utils.clip_grad_norm_(list(ref.values()), 1)
        topt.step(); opt.step()
    for n, p in m.live_parameters():
        close(p, ref[n], rtol=1e-5, atol=1e-7, what=n)


# (256, 384, 1024) .. (640, 1100, 328) run on the CTA-pair kernel (256x256 tiles; fn_tc_gemm2.cu), the others on 128x128 tiles
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 1024), (200, 342, 520), (1000, 72, 136), (64, 8, 8),
                                   (512, 512, 64), (300, 640, 200), (2100, 1024, 1544), (640, 1100, 328), (700, 342, 520), (256, 352, 200), (1000, 300, 72)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("c_bf16", [0, 1])
def test_tc_gemm_bf16(dev, M, N, K, a_mn, b_mn, c_bf16):
    """tcgen05 GEMM, every operand major-ness, ragged edges; reference = fp32 matmul of the same bf16 values."""
    from fadernets_b200._lib import LIB
    from fadernets_b200.ops import _p, _st
    r8 = lambda x: (x + 7) // 8 * 8
    bf = torch.bfloat16
    if not a_mn:
        Abuf = torch.zeros(M, r8(K), dtype=bf, device=dev); Abuf[:, :K] = rnd(M, K, seed=1, dev=dev).to(bf); A = Abuf[:, :K].float(); lda = r8(K)
    else:
        Abuf = torch.zeros(K, r8(M), dtype=bf, device=dev); Abuf[:, :M] = rnd(K, M, seed=1, dev=dev).to(bf); A = Abuf[:, :M].float().t(); lda = r8(M)
    if not b_mn:
        Bbuf = torch.zeros(N, r8(K), dtype=bf, device=dev); Bbuf[:, :K] = rnd(N, K, seed=2, dev=dev).to(bf); Bm = Bbuf[:, :K].float().t(); ldb = r8(K)
    else:
        Bbuf = torch.zeros(K, r8(N), dtype=bf, device=dev); Bbuf[:, :N] = rnd(K, N, seed=2, dev=dev).to(bf); Bm = Bbuf[:, :N].float(); ldb = r8(N)
    bias = rnd(N, seed=3, dev=dev)
    C0 = rnd(M, N + 5, seed=4, dev=dev)
    Cm = C0.to(bf) if c_bf16 else C0.clone()
    Cin = Cm.clone()
    LIB.call("fn_tc_gemm_bf16", _p(Abuf), lda, a_mn, _p(Bbuf), ldb, b_mn, _p(Cm), N + 5, c_bf16, _p(bias), M, N, K, 1, _st(Cm))
    torch.cuda.synchronize()
    ref = Cin.double()
    ref[:, :N] += A.double() @ Bm.double() + bias.double()
    if c_bf16:
        close(Cm[:, :N], ref[:, :N], rtol=1e-2, atol=1e-2, what="tc gemm (bf16 out)")
    else:
        close(Cm[:, :N], ref[:, :N], rtol=2e-5, atol=1e-4, what="tc gemm (fp32 out)")
    assert torch.equal(Cm[:, N:], Cin[:, N:]), "wrote outside the tile"


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core path
# ------------------------------------------------------------------------------------------------
def _ste_bf16(x):
    """bf16 rounding with a straight-through gradient (what the kernels do to MMA operands)."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def _torch_gru_bf16(gi, h0, w_hh, b_hh, reverse):
    """_torch_gru with the tensor-core path's operand rounding: bf16 W_hh and bf16 state INTO the product,
    fp32 (here fp64) everywhere else."""
    T, B, _ = gi.shape
    H = w_hh.shape[1]
    wq = _ste_bf16(w_hh)
    h = _ste_bf16(h0)                      # the initial-state slab is bf16
    hs = [None] * T
    for s in range(T):
        t = T - 1 - s if reverse else s
        gh = _ste_bf16(h) @ wq.t() + b_hh
        r = torch.sigmoid(gi[t, :, :H] + gh[:, :H])
        z = torch.sigmoid(gi[t, :, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[t, :, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        hs[t] = h
    return torch.stack(hs, 0)


@pytest.mark.parametrize("B,T,H,Vin,Zin", [(3, 5, 64, 7, 4), (70, 9, 64, 342, 12), (200, 6, 128, 16, 8),
                                           (130, 4, 256, 3, 8), (256, 3, 512, 342, 24), (5, 1, 64, 3, 8)])
def test_gru_group_bf16(dev, B, T, H, Vin, Zin):
    """tcgen05 GRU: three chains in one launch -- (emb, reverse, zero h0, final state) / (emb + z-projection,
    h0) / (dense bf16 input, h0 = xin[0]) -- forward values and every gradient against a torch fp64
    restatement with the same bf16 operand rounding.  Tolerances are bf16-level (saved gates and gate
    gradients are stored in bf16)."""
    from fadernets_b200.ops import ChainSpec
    from fadernets_b200.ops_bf16 import GruGroupBf16Fn
    bf = torch.bfloat16
    k = 1.0 / math.sqrt(H)
    def par(*shape, seed):
        return (rnd(*shape, seed=seed, dev=dev) * k).requires_grad_(True)
    ids = torch.randint(0, Vin, (T, B), generator=torch.Generator().manual_seed(5)).int().to(dev)
    wa = [par(3 * H, Vin, seed=10), par(3 * H, seed=11), par(3 * H, H, seed=12), par(3 * H, seed=13)]
    wb = [par(3 * H, Vin + Zin, seed=20), par(3 * H, seed=21), par(3 * H, H, seed=22), par(3 * H, seed=23)]
    zin = rnd(B, Zin, seed=24, dev=dev).requires_grad_(True)
    h0b = rnd(B, H, seed=25, dev=dev).requires_grad_(True)
    wc = [par(3 * H, H, seed=30), par(3 * H, seed=31), par(3 * H, H, seed=32), par(3 * H, seed=33)]
    xin = rnd(T, B, H, seed=34, dev=dev, scale=0.5).to(bf).requires_grad_(True)
    specs = [ChainSpec(emb_cols=(0, Vin), ids=ids, reverse=True, final=(0, 2)),
             ChainSpec(emb_cols=(0, Vin), ids=ids, z_cols=(Vin, Zin), h0="tensor", want_hs=True),
             ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)]
    fin, hs_b, hs_c = GruGroupBf16Fn.apply(specs, B, T, H, (H + 5,), *wa, *wb, zin, h0b, *wc, xin)
    assert hs_b.dtype == bf and hs_c.dtype == bf and fin.dtype == torch.float32
    go_f, go_b, go_c = rnd(B, H, seed=40, dev=dev), rnd(T, B, H, seed=41, dev=dev).to(bf), rnd(T, B, H, seed=42, dev=dev).to(bf)
    loss = (fin[:, 2:2 + H] * go_f).sum() + (hs_b.float() * go_b.float()).sum() + (hs_c.float() * go_c.float()).sum()
    leaves = wa + wb + [zin, h0b] + wc + [xin]
    grads = torch.autograd.grad(loss, leaves)
    torch.cuda.synchronize()

    D = [t.detach().double().requires_grad_(True) for t in leaves]
    a_wih, a_bih, a_whh, a_bhh, b_wih, b_bih, b_whh, b_bhh, zin_, h0b_, c_wih, c_bih, c_whh, c_bhh, xin_ = D
    idl = ids.long()
    gi_a = _ste_bf16(a_wih).t()[idl] + a_bih                        # the gathered table is stored in bf16
    ra = _torch_gru_bf16(gi_a, torch.zeros(B, H, dtype=torch.float64, device=dev), a_whh, a_bhh, True)
    gi_b = _ste_bf16(b_wih[:, :Vin]).t()[idl] + (zin_ @ b_wih[:, Vin:].t() + b_bih)[None]
    rb = _torch_gru_bf16(gi_b, h0b_, b_whh, b_bhh, False)
    gi_c = _ste_bf16(xin_ @ _ste_bf16(c_wih).t() + c_bih)          # the dense stream is stored in bf16
    rc = _torch_gru_bf16(gi_c, xin_[0], c_whh, c_bhh, False)
    close(fin[:, 2:2 + H], ra[0], rtol=2e-3, atol=2e-3, what="final state (reverse chain)")
    close(hs_b.float(), rb, rtol=1e-2, atol=5e-3, what="hs chain B")
    close(hs_c.float(), rc, rtol=1e-2, atol=5e-3, what="hs chain C")
    rloss = (ra[0] * go_f.double()).sum() + (rb * go_b.double()).sum() + (rc * go_c.double()).sum()
    rgrads = torch.autograd.grad(rloss, D)
    names = ["a_wih", "a_bih", "a_whh", "a_bhh", "b_wih", "b_bih", "b_whh", "b_bhh", "zin", "h0b", "c_wih", "c_bih",
             "c_whh", "c_bhh", "xin"]
    for nm, g, rg in zip(names, grads, rgrads):
        close(g.float(), rg, rtol=3e-2, atol=1e-3, what="grad " + nm)


def test_bf16_aux_kernels(dev):
    from fadernets_b200._lib import LIB
    from fadernets_b200.ops import _p, _st
    from fadernets_b200 import ops_bf16 as ob
    bf = torch.bfloat16
    # casts (contiguous, padded pitch, transposing)
    x = rnd(37, 50, seed=1, dev=dev)
    assert torch.equal(ob.cast_bf16(x, 37, 50, 50, 1), x.to(bf))
    p = ob.cast_bf16(x, 37, 50, 50, 1, ld_dst=56)
    assert torch.equal(p[:, :50], x.to(bf))
    assert torch.equal(ob.cast_bf16(x, 50, 37, 1, 50), x.t().to(bf))
    y = rnd(64, 128, seed=2, dev=dev)
    assert torch.equal(ob.cast_bf16(y, 64, 128, 128, 1), y.to(bf))
    # one-hot operand
    ids = torch.randint(0, 342, (1000,), generator=torch.Generator().manual_seed(3)).int().to(dev)
    oh = ob.onehot_bf16(ids, 342)
    assert oh.shape == (1000, 344)
    assert torch.equal(oh[:, :342].float(), torch.nn.functional.one_hot(ids.long(), 342).float()) and float(oh[:, 342:].abs().sum()) == 0
    oh3 = ob.onehot_bf16(ids % 3, 3)
    assert torch.equal(oh3[:, :3].float(), torch.nn.functional.one_hot((ids % 3).long(), 3).float())
    # time sums over dg [T][B][4H]
    T, B, H = 7, 5, 16
    dg = rnd(T, B, 4 * H, seed=4, dev=dev).to(bf)
    dproj = torch.empty(B, 3 * H, device=dev); dgh = torch.empty(B, 3 * H, device=dev)
    LIB.call("fn_time_sum_bf16", _p(dg), B, T, H, _p(dproj), _p(dgh), _st(dg))
    s = dg.float().sum(0)
    close(dproj, s[:, :3 * H], what="time sum dproj")
    close(dgh, torch.cat([s[:, :2 * H], s[:, 3 * H:]], 1), what="time sum dgh")
    # column sum
    xb = rnd(1000, 24, seed=5, dev=dev).to(bf)
    out = torch.ones(19, device=dev)
    ob.col_sum_bf16(xb, 24, 1000, 19, out, accumulate=True)
    close(out, xb.float()[:, :19].sum(0) + 1, rtol=1e-5, atol=1e-4, what="col sum bf16")
    # bf16 += fp32
    a = rnd(100, seed=6, dev=dev).to(bf); b = rnd(100, seed=7, dev=dev)
    a0 = a.clone()
    LIB.call("fn_add_f32_to_bf16", _p(a), _p(b), 100, _st(a))
    assert torch.equal(a, (a0.float() + b).to(bf))


@pytest.mark.parametrize("M,N,K", [(300, 342, 64), (1000, 3, 128), (257, 16, 256)])
def test_linear_bf16(dev, M, N, K):
    from fadernets_b200.ops_bf16 import linear_bf16
    bf = torch.bfloat16
    x = rnd(M, K, seed=1, dev=dev).to(bf).requires_grad_(True)
    w = (rnd(N, K, seed=2, dev=dev) / math.sqrt(K)).requires_grad_(True)
    b = rnd(N, seed=3, dev=dev).requires_grad_(True)
    y = linear_bf16(x, w, b)
    go = rnd(M, N, seed=4, dev=dev)
    gx, gw, gb = torch.autograd.grad((y * go).sum(), [x, w, b])
    xd, wd, bd = x.detach().double().requires_grad_(True), w.detach().to(bf).double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = xd @ wd.t() + bd
    close(y, yr, rtol=1e-4, atol=1e-4, what="linear bf16 fwd")
    gob = go.to(bf).double()
    rgx, rgw, rgb = torch.autograd.grad((yr * gob).sum(), [xd, wd, bd])
    assert gx.dtype == bf
    close(gx.float(), rgx, rtol=1e-2, atol=1e-2, what="dx"); close(gw, rgw, rtol=1e-4, atol=1e-3, what="dw")
    close(gb, rgb, rtol=1e-4, atol=1e-3, what="db")


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1)])
@pytest.mark.parametrize("c_bf16", [0, 1])
def test_tc_gemm_splitk(dev, a_mn, b_mn, c_bf16):
    """Split-K path (workspace partials + fixed-order reduce), ragged K split, bias + accumulate; and twice the
    same call gives bit-identical results (no float atomics)."""
    from fadernets_b200._lib import LIB
    from fadernets_b200.ops import _p, _st
    bf = torch.bfloat16
    _splitk_case(dev, a_mn, b_mn, c_bf16, 304, 200, 5000, 7)        # 128x128-tile kernel
    _splitk_case(dev, a_mn, b_mn, c_bf16, 520, 768, 5000, 7)        # CTA-pair kernel (ragged last K split, ragged M)
    _splitk_case(dev, a_mn, b_mn, c_bf16, 3072, 344, 9000, 4)       # 176-column pair tiles (vocabulary width)
    _splitk_case(dev, a_mn, b_mn, c_bf16, 512, 512, 640, 4)         # 10 K blocks over 4 splits: re-planned to non-empty splits


def _splitk_case(dev, a_mn, b_mn, c_bf16, M, N, K, splits):
    from fadernets_b200._lib import LIB
    from fadernets_b200.ops import _p, _st
    bf = torch.bfloat16                        # lda = M, ldb = N in the MN-major layouts: multiples of 8
    A = rnd(M, K, seed=1, dev=dev).to(bf); Bm = rnd(K, N, seed=2, dev=dev).to(bf)
    Abuf = A.t().contiguous() if a_mn else A.contiguous()              # a_mn: stored [K][M]
    Bbuf = Bm.contiguous() if b_mn else Bm.t().contiguous()            # b_mn: stored [K][N]; else [N][K]
    lda = M if a_mn else K
    ldb = N if b_mn else K
    bias = rnd(N, seed=3, dev=dev)
    C0 = rnd(M, N, seed=4, dev=dev)
    outs = []
    for rep in range(2):
        Cm = C0.to(bf) if c_bf16 else C0.clone()
        nb = LIB.call("fn_tc_gemm_splitk_ws_bytes", M, N, splits)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        LIB.call("fn_tc_gemm_bf16_splitk", _p(Abuf), lda, a_mn, _p(Bbuf), ldb, b_mn, _p(Cm), N, c_bf16, _p(bias), M, N, K, 1,
                 splits, _p(ws), nb, _st(Cm))
        torch.cuda.synchronize()
        outs.append(Cm.clone())
    assert torch.equal(outs[0], outs[1]), "split-K result is not run-to-run deterministic"
    ref = (C0.to(bf).double() if c_bf16 else C0.double()) + A.double() @ Bm.double() + bias.double()
    close(outs[0], ref, rtol=1e-2 if c_bf16 else 2e-5, atol=5e-2 if c_bf16 else 2e-3, what="split-K gemm")


def test_index_validation(dev):
    """Out-of-range token / target / label indices: the reference raises (F.nll_loss, nn.Embedding); here they are clamped
    on the device (no out-of-bounds access), counted, and raised as IndexError at the next synchronisation."""
    from fadernets_b200 import ops
    ops.raise_on_bad_index()                                   # clear
    ids = torch.tensor([[0, 5, 341, 7], [1, 2, 3, 4]], device=dev)
    tm = ops.ids_to_tm(ids, dims=342)
    ops.raise_on_bad_index()
    assert tm.t().tolist() == ids.tolist()
    bad = ids.clone(); bad[0, 1] = 999; bad[1, 3] = -3
    tm = ops.ids_to_tm(bad, dims=342)
    assert tm.t().tolist() == [[0, 341, 341, 7], [1, 2, 3, 0]]
    with pytest.raises(IndexError):
        ops.raise_on_bad_index()
    ops.raise_on_bad_index()                                   # the counter was reset
    logp = torch.log_softmax(rnd(6, 5, seed=1, dev=dev), -1).requires_grad_(True)
    tgt = torch.tensor([0, 4, 2, 7, 1, -1], device=dev)
    loss = ops.NllMeanFn.apply(logp, tgt)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(logp.grad).all()
    with pytest.raises(IndexError):
        ops.raise_on_bad_index()
    # a one-hot row without a maximum still yields a valid id
    oh = torch.full((1, 2, 4), float("nan"), device=dev)
    assert ops.onehot_to_ids_tm(oh).flatten().tolist() == [0, 0]
