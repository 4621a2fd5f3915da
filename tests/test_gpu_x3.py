"""GPU parity of the bf16x3 mode's building blocks (fn_split_bf16, fn_tc_gemm_bf16x3, fn_gru_seq_*_bf16x3, LinearX3Fn):
tensor-core products over hi / lo bf16 planes against torch fp64 on the same fp32 inputs, at fp32-level tolerances
(the mode claims north_star's 1e-3 relative bar; observed errors are ~1e-5)."""
import math

import pytest
import torch

from test_gpu_ops import _torch_gru, close, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_split_planes(dev):
    from fadernets_b200 import ops_x3 as ox
    x = rnd(37, 50, seed=1, dev=dev)
    s, ld, lo = ox.split_bf16(x, 37, 50, 50, 1)
    assert (ld, lo) == (112, 56) and s.shape == (37, 112)
    hi, lw = s[:, :50].float(), s[:, 56:106].float()
    assert torch.equal(hi, x.to(torch.bfloat16).float())
    assert float((hi + lw - x).abs().max()) <= 2.0 ** -16 * float(x.abs().max())
    st, ld, lo = ox.split_bf16(x, 50, 37, 1, 50)                      # transposing
    assert torch.equal(st[:, :37].float(), x.t().to(torch.bfloat16).float())
    assert float((st[:, :37].float() + st[:, lo:lo + 37].float() - x.t()).abs().max()) <= 2.0 ** -16 * float(x.abs().max())
    w = rnd(48, 64, seed=2, dev=dev)
    t3, ld, lo = ox.split_bf16(w, 48, 64, 64, 1, triple=True)         # [hi | hi | lo]
    assert (ld, lo) == (192, 128)
    assert torch.equal(t3[:, :64], t3[:, 64:128]) and torch.equal(t3[:, :64].float(), w.to(torch.bfloat16).float())
    assert float((t3[:, :64].float() + t3[:, 128:].float() - w).abs().max()) <= 2.0 ** -16 * float(w.abs().max())


@pytest.mark.parametrize("M,N,K", [(5, 7, 8), (130, 342, 72), (257, 129, 520), (300, 1536, 64), (96, 64, 16384), (1000, 3, 512)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_tc_gemm_x3(dev, M, N, K, a_mn, b_mn):
    """(A_hi + A_lo)(B_hi + B_lo) - A_lo B_lo with fp32 accumulation vs fp64: error ~2^-16 of |A||B| per term."""
    from fadernets_b200 import ops_x3 as ox
    A = rnd(M, K, seed=1, dev=dev)
    Bm = rnd(N, K, seed=2, dev=dev)
    bias = rnd(N, seed=3, dev=dev)
    # K-major operand: [rows][2*r8(K)] planes side by side;  MN-major: memory [K][2*r8(rows)]
    As = ox.split_bf16(A, M, K, K, 1) if not a_mn else ox.split_bf16(A, K, M, 1, K)
    Bs = ox.split_bf16(Bm, N, K, K, 1) if not b_mn else ox.split_bf16(Bm, K, N, 1, K)
    C0 = rnd(M, N + 5, seed=4, dev=dev)
    Cm = C0.clone()
    ox.tc_gemm_x3(As[0], 0, As[1], As[2], a_mn, Bs[0], 0, Bs[1], Bs[2], b_mn, Cm, 0, N + 5, bias, M, N, K, accumulate=True)
    ref = C0.double().clone()
    ref[:, :N] += A.double() @ Bm.double().t() + bias.double()
    scale = float((A.double().abs() @ Bm.double().abs().t()).max())
    err = float((Cm.double() - ref)[:, :N].abs().max())
    assert err <= 4e-5 * scale + 1e-5, (err, scale)
    assert torch.equal(Cm[:, N:], C0[:, N:]), "wrote outside the tile"
    # an exact operand (one plane): B rounded to bf16 first
    Bq = Bm.to(torch.bfloat16)
    Bq2 = Bq if not b_mn else Bq.t().contiguous()
    if (Bq2.shape[1] % 8) == 0:
        C2 = torch.empty(M, N, device=dev)
        ox.tc_gemm_x3(As[0], 0, As[1], As[2], a_mn, Bq2, 0, Bq2.shape[1], 0, b_mn, C2, 0, N, None, M, N, K)
        ref2 = A.double() @ Bq.double().t()
        assert float((C2.double() - ref2).abs().max()) <= 4e-5 * scale + 1e-5


@pytest.mark.parametrize("M,N,K", [(70, 342, 64), (1000, 3, 128), (513, 16, 256)])
def test_linear_x3(dev, M, N, K):
    from fadernets_b200 import ops_x3 as ox
    xf = rnd(M, K, seed=1, dev=dev)
    xs = ox.split_bf16(xf, M, K, K, 1)[0].requires_grad_(True)
    w = (rnd(N, K, seed=2, dev=dev) / math.sqrt(K)).requires_grad_(True)
    b = rnd(N, seed=3, dev=dev).requires_grad_(True)
    go = rnd(M, N, seed=4, dev=dev)
    y = ox.linear_x3(xs, w, b)
    gx, gw, gb = torch.autograd.grad((y * go).sum(), (xs, w, b))
    xv = (xs[:, :K].float() + xs[:, K:].float()).double().requires_grad_(True)        # the value the kernel sees
    wd, bd = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = xv @ wd.t() + bd
    rx, rw, rb = torch.autograd.grad((yr * go.double()).sum(), (xv, wd, bd))
    close(y, yr, rtol=1e-4, atol=1e-5, what="y")
    close(ox.from_split_grad(gx), rx, rtol=1e-4, atol=1e-5, what="dx")
    close(gw, rw, rtol=1e-4, atol=1e-5, what="dW")
    close(gb, rb, rtol=1e-4, atol=1e-5, what="db")


@pytest.mark.parametrize("B,T,H,Vin,Zin", [(3, 5, 64, 7, 4), (70, 9, 64, 342, 12), (200, 6, 128, 16, 8), (130, 4, 256, 3, 8),
                                           (64, 24, 512, 342, 24), (5, 1, 64, 3, 8)])
def test_gru_group_x3(dev, B, T, H, Vin, Zin):
    """tcgen05 GRU in bf16x3 mode: the three chain kinds of test_gru_group_bf16 -- forward values and EVERY gradient against
    the plain fp64 GRU (no operand rounding in the restatement: the mode claims fp32-level results)."""
    from fadernets_b200 import ops_x3 as ox
    from fadernets_b200.ops import ChainSpec
    k = 1.0 / math.sqrt(H)

    def par(*shape, seed):
        return (rnd(*shape, seed=seed, dev=dev) * k).requires_grad_(True)
    ids = torch.randint(0, Vin, (T, B), generator=torch.Generator().manual_seed(5)).int().to(dev)
    wa = [par(3 * H, Vin, seed=10), par(3 * H, seed=11), par(3 * H, H, seed=12), par(3 * H, seed=13)]
    wb = [par(3 * H, Vin + Zin, seed=20), par(3 * H, seed=21), par(3 * H, H, seed=22), par(3 * H, seed=23)]
    zin = rnd(B, Zin, seed=24, dev=dev).requires_grad_(True)
    h0b = rnd(B, H, seed=25, dev=dev).requires_grad_(True)
    wc = [par(3 * H, H, seed=30), par(3 * H, seed=31), par(3 * H, H, seed=32), par(3 * H, seed=33)]
    xf = rnd(T, B, H, seed=34, dev=dev, scale=0.5)
    xin = ox.split_bf16(xf, T * B, H, H, 1)[0].view(T, B, 2 * H).requires_grad_(True)
    specs = [ChainSpec(emb_cols=(0, Vin), ids=ids, reverse=True, final=(0, 2)),
             ChainSpec(emb_cols=(0, Vin), ids=ids, z_cols=(Vin, Zin), h0="tensor", want_hs=True),
             ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)]
    fin, hs_b, hs_c = ox.GruGroupX3Fn.apply(specs, B, T, H, (H + 5,), *wa, *wb, zin, h0b, *wc, xin)
    assert hs_b.shape == (T, B, 2 * H) and hs_b.dtype == torch.bfloat16 and fin.dtype == torch.float32
    val = lambda s: s[..., :H].float() + s[..., H:].float()
    go_f, go_b, go_c = rnd(B, H, seed=40, dev=dev), rnd(T, B, H, seed=41, dev=dev), rnd(T, B, H, seed=42, dev=dev)
    leaves = wa + wb + [zin, h0b] + wc + [xin]
    # the gradient wrt a split tensor's value is handed over in the bf16 view of an fp32 buffer
    grads = torch.autograd.grad((fin, hs_b, hs_c), leaves, grad_outputs=(torch.nn.functional.pad(go_f, (2, 3)),
                                ox.as_split_grad(go_b.contiguous()), ox.as_split_grad(go_c.contiguous())))
    torch.cuda.synchronize()

    D = [t.detach().double().requires_grad_(True) for t in leaves[:-1]] + [val(xin.detach()).double().requires_grad_(True)]
    a_wih, a_bih, a_whh, a_bhh, b_wih, b_bih, b_whh, b_bhh, zin_, h0b_, c_wih, c_bih, c_whh, c_bhh, xin_ = D
    idl = ids.long()
    ra = _torch_gru(a_wih.t()[idl] + a_bih, torch.zeros(B, H, dtype=torch.float64, device=dev), a_whh, a_bhh, True)
    rb = _torch_gru(b_wih[:, :Vin].t()[idl] + (zin_ @ b_wih[:, Vin:].t() + b_bih)[None], h0b_, b_whh, b_bhh, False)
    rc = _torch_gru(xin_ @ c_wih.t() + c_bih, xin_[0], c_whh, c_bhh, False)
    close(fin[:, 2:2 + H], ra[0], rtol=1e-4, atol=2e-5, what="final state (reverse chain)")
    close(val(hs_b), rb, rtol=1e-4, atol=2e-5, what="hs chain B")
    close(val(hs_c), rc, rtol=1e-4, atol=2e-5, what="hs chain C")
    rloss = (ra[0] * go_f.double()).sum() + (rb * go_b.double()).sum() + (rc * go_c.double()).sum()
    rgrads = torch.autograd.grad(rloss, D)
    names = ["a_wih", "a_bih", "a_whh", "a_bhh", "b_wih", "b_bih", "b_whh", "b_bhh", "zin", "h0b", "c_wih", "c_bih",
             "c_whh", "c_bhh", "xin"]
    for nm, g, rg in zip(names, grads, rgrads):
        g = ox.from_split_grad(g) if nm == "xin" else g
        close(g, rg, rtol=3e-4, atol=2e-5, what="grad " + nm)
