"""Parity of the CTA-pair GRU kernels (fn_gru_tc2.cu: tcgen05.mma cta_group::2, the instantiations bench.py times at
config 3) against a torch fp64 restatement of the GRU equations (torch nn/modules/rnn.py GRU docstring; reference
call sites gmm_model.py:84,89,109,114,133,136) with the SAME bf16 operand rounding.

Covers what round 1 left unchecked: H = 1024 (partially resident weights + the streamed weight ring), H = 512 / 128
(everything resident), ragged second batch tile (B = 130, 200), reverse chains, token / projection / dense inputs,
many steps (ring wrap-around, barrier parities) and the full T = 512 sequence."""
import math

import pytest
import torch

from test_gpu_ops import _ste_bf16, _torch_gru_bf16, close, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,T,H,Vin,Zin", [(256, 16, 1024, 342, 24), (130, 48, 1024, 16, 8), (256, 40, 512, 342, 24),
                                           (200, 33, 128, 3, 8), (129, 7, 64, 16, 8)])
def test_pair_kernels_three_chains(dev, B, T, H, Vin, Zin):
    """Three chains in one launch -- (token gather, reverse, zero h0, final state) / (token gather + z projection, h0)
    / (dense bf16 input, h0 = xin[0]) -- forward states and every gradient."""
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
    go_f, go_b, go_c = rnd(B, H, seed=40, dev=dev), rnd(T, B, H, seed=41, dev=dev).to(bf), rnd(T, B, H, seed=42, dev=dev).to(bf)
    loss = (fin[:, 2:2 + H] * go_f).sum() + (hs_b.float() * go_b.float()).sum() + (hs_c.float() * go_c.float()).sum()
    leaves = wa + wb + [zin, h0b] + wc + [xin]
    grads = torch.autograd.grad(loss, leaves)
    torch.cuda.synchronize()

    D = [t.detach().double().requires_grad_(True) for t in leaves]
    a_wih, a_bih, a_whh, a_bhh, b_wih, b_bih, b_whh, b_bhh, zin_, h0b_, c_wih, c_bih, c_whh, c_bhh, xin_ = D
    idl = ids.long()
    gi_a = _ste_bf16(a_wih).t()[idl] + a_bih
    ra = _torch_gru_bf16(gi_a, torch.zeros(B, H, dtype=torch.float64, device=dev), a_whh, a_bhh, True)
    gi_b = _ste_bf16(b_wih[:, :Vin]).t()[idl] + (zin_ @ b_wih[:, Vin:].t() + b_bih)[None]
    rb = _torch_gru_bf16(gi_b, h0b_, b_whh, b_bhh, False)
    gi_c = _ste_bf16(xin_ @ _ste_bf16(c_wih).t() + c_bih)
    rc = _torch_gru_bf16(gi_c, xin_[0], c_whh, c_bhh, False)
    close(fin[:, 2:2 + H], ra[0], rtol=3e-3, atol=3e-3, what="final state (reverse chain)")
    close(hs_b.float(), rb, rtol=1e-2, atol=5e-3, what="hs chain B")
    close(hs_c.float(), rc, rtol=1e-2, atol=5e-3, what="hs chain C")
    rloss = (ra[0] * go_f.double()).sum() + (rb * go_b.double()).sum() + (rc * go_c.double()).sum()
    rgrads = torch.autograd.grad(rloss, D)
    names = ["a_wih", "a_bih", "a_whh", "a_bhh", "b_wih", "b_bih", "b_whh", "b_bhh", "zin", "h0b", "c_wih", "c_bih",
             "c_whh", "c_bhh", "xin"]
    for nm, g, rg in zip(names, grads, rgrads):
        close(g.float(), rg, rtol=3e-2, atol=1e-3 * max(1.0, T / 8), what="grad " + nm)


@pytest.mark.parametrize("B,T,H", [(256, 512, 1024), (256, 256, 512)])
def test_pair_forward_full_length(dev, B, T, H):
    """One token-input chain over the FULL sequence length of configs 3 / 2 (ring wrap-around, barrier phases, the
    counters after hundreds of steps): every state against the restatement; forward again gives identical bits."""
    from fadernets_b200.ops import ChainSpec
    from fadernets_b200.ops_bf16 import GruGroupBf16Fn
    V = 342
    k = 1.0 / math.sqrt(H)
    ids = torch.randint(0, V, (T, B), generator=torch.Generator().manual_seed(7)).int().to(dev)
    w = [rnd(3 * H, V, seed=1, dev=dev) * k, rnd(3 * H, seed=2, dev=dev) * k, rnd(3 * H, H, seed=3, dev=dev) * k,
         rnd(3 * H, seed=4, dev=dev) * k]
    h0 = rnd(B, H, seed=5, dev=dev)
    specs = [ChainSpec(emb_cols=(0, V), ids=ids, h0="tensor", want_hs=True)]
    with torch.no_grad():
        (hs,) = GruGroupBf16Fn.apply(specs, B, T, H, (), *w, h0)
        (hs2,) = GruGroupBf16Fn.apply(specs, B, T, H, (), *w, h0)
        torch.cuda.synchronize()
        assert torch.equal(hs, hs2), "forward is not run-to-run deterministic"
        wd = [t.double() for t in w]
        gi = _ste_bf16(wd[0]).t()[ids.long()] + wd[1]
        ref = _torch_gru_bf16(gi, h0.double(), wd[2], wd[3], False)
    err = (hs.double() - ref).abs()
    # bf16 state storage: 2^-9 relative per step on |h| <= 1, contracted by the gates; no growth with T
    assert float(err.max()) < 2.5e-2, float(err.max())
    assert float(err[T // 2:].mean()) < 2e-3, float(err[T // 2:].mean())
