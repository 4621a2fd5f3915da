"""GPU parity of the bf16 tensor-core mode (model.set_precision("bf16"); BASELINE configs 3-5) against the
fp32 CPU oracle.  The arithmetic type differs from the reference's fp32 here by design (bf16 MMA operands,
fp32 accumulation; bf16 storage of states / gates / gate gradients), so the tolerance is the bf16 one and
is written out below: 2e-2 relative on losses and outputs, 8e-2 of the largest element on gradients.
Arg-max of the mixture responsibilities stays bit-exact (latent block is fp32)."""
import numpy as np
import pytest
import torch

from oracle import fader_oracle as fo

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-2
GRAD_RTOL = 8e-2


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _model(variant, H, Z, K, w, dev):
    import fadernets_b200 as fn
    if variant == "gmvae":
        m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    else:
        m = fn.MusicAttrRegVAE(342, 3, 16, 24, H, Z, 32)
    m.load_state_dict(w)
    return m.to(dev).train().set_precision("bf16")


@pytest.mark.parametrize("variant,H,Z,K,B,T", [("gmvae", 64, 32, 2, 70, 33), ("vae", 128, 16, 0, 9, 40),
                                               ("gmvae", 256, 128, 2, 200, 24),
                                               # the H = 1024 instantiations bench.py times (CTA-pair GRU kernels with streamed
                                               # weights, 16-segment wavefront needs T >= 512: here the 1-2 segment forms)
                                               ("gmvae", 1024, 128, 2, 256, 16), ("gmvae", 1024, 128, 2, 130, 48)])
def test_bf16_train_step_vs_oracle(dev, variant, H, Z, K, B, T):
    import fadernets_b200 as fn
    from fadernets_b200 import trainer, trainer_gmm
    w = fo.init_weights(H, Z, variant, max(K, 1), seed=5)
    model = _model(variant, H, Z, K, w, dev)
    opt = fn.FusedAdam(model, lr=1e-3)
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=6, pad_tail=True)
    g = torch.Generator().manual_seed(8)
    er, en = torch.randn(B, Z, generator=g), torch.randn(B, Z, generator=g)
    scal, grads, res = fo.loss_and_grads(w, variant, (d, r, n, c, rd, nd), er, en, 20000, 0.2)
    it = iter((er, en))
    model._draw_eps = lambda B_, Z_, d_: next(it).to(d_)
    model.host_rng = False
    dd, rr, nn_, cc = d.to(dev), r.to(dev), n.to(dev), c.to(dev)
    opt.zero_grad()
    if variant == "gmvae":
        trainer_gmm.configure(model, opt, {"beta": 0.2})
        loss, terms, l_r, l_n = trainer_gmm._forward_losses(20000, dd, rr, nn_, dd, rr, nn_, cc, rd, nd, False, None)
        names = ("loss", "CE_X", "CE_R", "CE_N", "kld_lat_r", "kld_lat_n", "kld_cls_r", "kld_cls_n")
        for nm, tv in zip(names, terms):
            e = float(scal[nm])
            assert abs(float(tv) - e) <= LOSS_RTOL * max(1.0, abs(e)), (nm, float(tv), e)
    else:
        trainer.configure(model, opt, {"beta": 0.2}, step_=20000)
        loss, *_ = trainer._forward_losses(dd, rr, nn_, dd, rr, nn_, cc, rd, nd)
    e = float(scal["loss"])
    assert abs(float(loss) - e) <= LOSS_RTOL * max(1.0, abs(e)), (float(loss), e)
    loss.backward()
    torch.cuda.synchronize()
    bad = []
    params = dict(model.named_parameters())
    for k, ref in grads.items():
        if k in ("linear_out_r.bias", "linear_out_n.bias"):
            continue      # mathematically zero (time-axis log-softmax is shift invariant): rounding noise only
        got = params[k].grad.cpu()
        assert torch.isfinite(got).all(), k
        scale = max(float(ref.abs().max()), 1e-6)
        err = float((got - ref).abs().max())
        if err > GRAD_RTOL * (1.5 if T >= 48 else 1.0) * scale + 1e-6:      # (bf16 BPTT noise grows with T: same figure with either kernel generation)
            bad.append((k, round(err / scale, 4)))
    assert not bad, bad
    # cosine similarity of the whole gradient: direction must be essentially the fp32 one
    a = torch.cat([params[k].grad.cpu().reshape(-1) for k in grads])
    b = torch.cat([grads[k].reshape(-1) for k in grads])
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 0.999, cos


def test_bf16_training_reduces_loss(dev):
    """Ten optimiser steps in bf16 mode on one batch: the loss goes down and stays finite."""
    import fadernets_b200 as fn
    from fadernets_b200 import trainer_gmm
    H, Z, K, B, T = 128, 32, 2, 130, 16
    w = fo.init_weights(H, Z, "gmvae", K, seed=1)
    model = _model("gmvae", H, Z, K, w, dev)
    opt = fn.FusedAdam(model, lr=2e-3)
    trainer_gmm.configure(model, opt, {"beta": 0.2, "lr": 2e-3})
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=2)
    dd, rr, nn_, cc = d.to(dev), r.to(dev), n.to(dev), c.to(dev)
    torch.manual_seed(0)
    losses, step = [], 20000
    for _ in range(10):
        step, out = trainer_gmm.train(step, dd, rr, nn_, dd, rr, nn_, cc, rd, nd)
        losses.append(out[0])
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses


def test_bf16_greedy_decode_tracks_fp32(dev):
    """Greedy decode on the tensor-core kernels: log-probs stay within bf16 tolerance of the fp32 path while both
    follow the same token prefix, and most tokens agree (arg-max near-ties may flip under bf16 rounding)."""
    H, Z, K, B, steps = 128, 32, 2, 70, 12
    w = fo.init_weights(H, Z, "gmvae", K, seed=3)
    m32 = _model("gmvae", H, Z, K, w, dev).set_precision("f32").eval()
    m16 = _model("gmvae", H, Z, K, w, dev).eval()
    g = torch.Generator().manual_seed(4)
    zc = torch.randn(B, 2 * Z + 24, generator=g).to(dev)
    lp32, t32 = m32.decode_greedy(zc, steps)
    lp16, t16 = m16.decode_greedy(zc, steps)
    assert lp16.shape == lp32.shape and t16.shape == t32.shape
    assert torch.isfinite(lp16).all()
    # first step has identical inputs: tight comparison
    assert float((lp16[:, 0] - lp32[:, 0]).abs().max()) < 5e-2
    same_prefix = (t16 == t32).cumprod(1).bool()
    agree = float(same_prefix.float().mean())
    assert agree > 0.7, agree
    out = m16.global_decoder(zc, steps)
    # tokens are the first arg-max of the logits; log-softmax can merge two logits one ulp apart
    assert float((out.argmax(-1) == t16).float().mean()) > 0.999


@pytest.mark.parametrize("H,B,steps", [(128, 70, 12), (256, 200, 9), (64, 3, 5)])
def test_persistent_decode_matches_stepwise(dev, H, B, steps):
    """The one-kernel greedy decode (fn_decode_greedy_bf16) against the same loop issued as per-step launches of the
    already verified gate-block / GEMM kernels: same arithmetic, so the token streams agree and the log-probs match."""
    Z, K = 32, 2
    w = fo.init_weights(H, Z, "gmvae", K, seed=7)
    m = _model("gmvae", H, Z, K, w, dev).eval()
    g = torch.Generator().manual_seed(9)
    zc = torch.randn(B, 2 * Z + 24, generator=g).to(dev)
    m.decode_persistent = True
    lp_p, t_p = m.decode_greedy(zc, steps)
    assert m._decode_plans and all(p.persistent for p in m._decode_plans.values())
    m.decode_persistent = False
    lp_s, t_s = m.decode_greedy(zc, steps)
    same_prefix = (t_p == t_s).cumprod(1).bool()
    # (the two paths accumulate K in a different order: a near-tie can flip and that sequence then diverges)
    assert float(same_prefix.float().mean()) > 0.9, float(same_prefix.float().mean())
    assert torch.equal(t_p[:, 0], t_s[:, 0])
    mask = same_prefix.unsqueeze(-1).expand_as(lp_p)
    assert float(((lp_p - lp_s).abs() * mask).max()) < 2e-2
    assert torch.allclose(lp_p.exp().sum(-1), torch.ones(B, steps, device=dev), atol=1e-4)
    # token-only variant returns the same tokens
    _, t_only = (m.__setattr__("decode_persistent", True) or m).decode_greedy(zc, steps, return_logp=False)
    assert torch.equal(t_only, t_p)


def _bf(x):
    return x.to(torch.bfloat16).double()


def _replay_decode(m, zc, tokens):
    """fp64 restatement of the eval-mode global_decoder (gmm_model.py:119-149) with the tensor-core path's operand
    rounding (bf16 weights / states / embedding rows / cell-2 input projection), TEACHER-FORCED with the kernel's own
    tokens so that the restated states never diverge from the kernel's: returns the logits of every step."""
    H, V = m._dims["H"], m._dims["V"]
    c1, c2, lo = m.grucell_g, m.grucell_g_2, m.linear_out_g
    D = lambda t: t.detach().double()
    z = D(zc)
    w_ih1, w_hh1 = D(c1.weight_ih), _bf(c1.weight_hh)
    emb1 = _bf(c1.weight_ih[:, :V].t())                                          # gathered table is stored in bf16
    proj1 = z @ w_ih1[:, V:].t() + D(c1.bias_ih)
    w_ih2, w_hh2, w_o = _bf(c2.weight_ih), _bf(c2.weight_hh), _bf(lo.weight)
    h1 = _bf(z @ D(m.linear_init_global.weight).t() + D(m.linear_init_global.bias))   # the initial-state slab is bf16
    h2 = None
    B, steps = tokens.shape
    tok = torch.full((B,), V - 1, dtype=torch.long, device=zc.device)
    out = []

    def cell(gi, h, w_hh, b_hh):
        gh = _bf(h) @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H]); zt = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        return (1 - zt) * n + zt * h

    for i in range(steps):
        h1 = cell(emb1[tok] + proj1, h1, w_hh1, D(c1.bias_hh))
        gi2 = _bf(_bf(h1) @ w_ih2.t() + D(c2.bias_ih))                           # cell-2 input projection is stored in bf16
        h2 = cell(gi2, h1 if i == 0 else h2, w_hh2, D(c2.bias_hh))              # step 0: hx[1] <- the new hx[0]
        out.append(_bf(h2) @ w_o.t() + D(lo.bias))
        tok = tokens[:, i]
    return torch.stack(out, 1)


@pytest.mark.parametrize("H,B,steps", [(1024, 256, 24), (256, 130, 40), (128, 70, 12)])
def test_bf16_greedy_tokens_are_argmax_of_rounded_restatement(dev, H, B, steps):
    """north_star: 'bit-exact for argmax token indices'.  Each token the one-kernel greedy decode emits must be THE
    arg-max of the restated logits (same bf16 operand rounding, fp64 accumulation) computed from the kernel's own token
    prefix; the only slack is a near-tie inside fp32 accumulation-order noise (top-2 logit gap < 2e-3), which must be
    rare (< 0.5 % of the positions)."""
    Z, K = 32, 2
    w = fo.init_weights(H, Z, "gmvae", K, seed=11)
    m = _model("gmvae", H, Z, K, w, dev).eval()
    zc = torch.randn(B, 2 * Z + 24, generator=torch.Generator().manual_seed(12)).to(dev)
    lp, toks = m.decode_greedy(zc, steps)
    ref = _replay_decode(m, zc, toks)
    top = ref.max(-1).values
    chosen = ref.gather(-1, toks.unsqueeze(-1)).squeeze(-1)
    gap = (top - chosen)
    assert float(gap.max()) < 2e-3, float(gap.max())
    exact = float((ref.argmax(-1) == toks).float().mean())
    assert exact > 0.995, exact                                # (the differing positions are the near-ties bounded above)
    # and the returned log-probs are the restated log-softmax
    assert float((lp.double() - torch.log_softmax(ref, -1)).abs().max()) < 3e-2
