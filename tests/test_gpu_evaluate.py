"""Decode drivers: clean_output, repar, batched shift() and arousal transfer."""
import numpy as np
import pytest
import torch

from oracle import fader_oracle as fo

pytestmark = pytest.mark.gpu


def _ref_clean(tokens_row):
    """test_class.py:44-50 restated on a numpy token row."""
    recon = np.trim_zeros(np.asarray(tokens_row))
    if 1 in recon:
        last_idx = np.argwhere(recon == 1)[0][0]
        recon = recon[:last_idx]
    return recon


def test_clean_tokens_matches_reference_rule(lib):
    from fadernets_b200 import evaluate as E
    dev = torch.device("cuda:0")
    g = np.random.default_rng(0)
    rows = [g.integers(0, 6, size=37) for _ in range(200)]
    rows += [np.zeros(37, dtype=np.int64), np.ones(37, dtype=np.int64), np.array([0] * 5 + [7, 9, 1, 4] + [0] * 28),
             np.array([3] * 37), np.array([0] * 36 + [5]), np.array([1] + [0] * 36)]
    toks = torch.from_numpy(np.stack(rows).astype(np.int64)).to(dev)
    got = E.clean_outputs(toks)
    for r, g_ in zip(rows, got):
        assert np.array_equal(g_, _ref_clean(r)), (r, g_)
    # single-sequence API on log-probs
    lp = torch.full((1, 6, 342), -10.0, device=dev)
    for i, t in enumerate([0, 0, 5, 7, 1, 3]):
        lp[0, i, t] = 0.0
    assert E.clean_output(lp).tolist() == [5, 7]


def test_shift_and_arousal_transfer_batched(lib):
    import fadernets_b200 as fn
    from fadernets_b200 import evaluate as E
    dev = torch.device("cuda:0")
    H, Z, K, B, T = 32, 16, 2, 6, 10
    w = fo.init_weights(H, Z, "gmvae", K, seed=2)
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    m.load_state_dict(w)
    m = m.to(dev).train()
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=3)
    torch.manual_seed(1)
    out, z0 = E.shift(m, d.to(dev), r.to(dev), n.to(dev), c, target_z_value=0.7, attr="note", steps=9)
    assert out.shape == (B, 9, 342) and z0.shape == (B,) and not m.training          # shift() leaves the model in eval
    assert torch.allclose(out.exp().sum(-1), torch.ones(B, 9, device=dev), atol=1e-4)
    # every row decodes exactly like a batch-1 call with the same latent (the reference's per-sample loop)
    torch.manual_seed(1)
    m.train()
    out1, _ = E.shift(m, d[:1].to(dev), r[:1].to(dev), n[:1].to(dev), c[:1], target_z_value=0.7, attr="note", steps=9)
    assert out1.shape == (1, 9, 342)
    toks = E.arousal_transfer(m, d.to(dev), c, lam=0.5, steps=12, sample=False)
    assert toks.shape == (B, 12) and toks.dtype == torch.int64
    toks2 = E.arousal_transfer(m, d.to(dev), c, lam=0.5, steps=12, sample=False)
    assert torch.equal(toks, toks2)
    z = torch.randn(B, Z, generator=torch.Generator().manual_seed(0)).to(dev)
    torch.manual_seed(3)
    a = E.repar(z, z.abs())
    torch.manual_seed(3)
    eps = torch.distributions.Normal(0, 1).sample(sample_shape=z.size()).to(dev)
    assert torch.equal(a, z + z.abs() * eps)


def test_shift_against_oracle(lib):
    """RhythmEvaluator.shift / NoteEvaluator.shift (test_class.py:233-254, 282-303): forward -> re-draw z from the
    returned distributions -> overwrite latent dim 0 -> eval-mode greedy decode, against the CPU oracle replaying the same
    CPU-generator draws (forward: eps_r, eps_n, T coin flips; then the two repar draws): tokens bit-exact, log-probs 1e-3."""
    import fadernets_b200 as fn
    from fadernets_b200 import evaluate as E
    dev = torch.device("cuda:0")
    H, Z, K, B, T, steps = 32, 16, 2, 5, 10, 14
    w = fo.init_weights(H, Z, "gmvae", K, seed=2)
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=3)
    for attr, target in (("rhythm", 0.7), ("note", -1.3)):
        m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
        m.load_state_dict(w)
        m = m.to(dev).train()
        torch.manual_seed(11)
        out, z0 = E.shift(m, d.to(dev), r.to(dev), n.to(dev), c, target_z_value=target, attr=attr, steps=steps)
        # ---- oracle with the same draws
        torch.manual_seed(11)
        er, en = fo.draw_eps(B, Z, T)
        res = fo.forward(w, "gmvae", d, r, n, c, er, en)
        e2r = torch.distributions.Normal(0, 1).sample(sample_shape=res["scale_r"].size())
        e2n = torch.distributions.Normal(0, 1).sample(sample_shape=res["scale_n"].size())
        z_r, z_n = res["mu_r"] + res["scale_r"] * e2r, res["mu_n"] + res["scale_n"] * e2n
        z_sel = z_r if attr == "rhythm" else z_n
        z0_ref = z_sel[:, 0].clone()
        z_sel[:, 0] = target
        lp_ref, tok_ref = fo.global_decoder(w, torch.cat([z_r, z_n, c], 1), steps)
        assert torch.allclose(z0.cpu(), z0_ref, rtol=1e-4, atol=1e-5)
        assert torch.equal(out.argmax(-1).cpu(), tok_ref), attr
        assert float((out.cpu() - lp_ref).abs().max()) < 1e-3 * max(1.0, float(lp_ref.abs().max()))


def test_arousal_transfer_against_oracle(lib):
    """arousal_transfer.ipynb cells 11/15/17: shift vectors mu_lookup(1) - mu_lookup(0), encode, z = mean + lam * shift,
    greedy decode: tokens bit-exact against the CPU oracle (fp32 path)."""
    import fadernets_b200 as fn
    from fadernets_b200 import evaluate as E
    dev = torch.device("cuda:0")
    H, Z, K, B, T, steps = 32, 16, 2, 6, 12, 20
    w = fo.init_weights(H, Z, "gmvae", K, seed=4)
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    m.load_state_dict(w)
    m = m.to(dev)
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=5)
    toks = E.arousal_transfer(m, d.to(dev), c, lam=0.5, steps=steps, sample=False)
    mu_r, s_r, mu_n, s_n = fo.encode(w, d)
    sr = w["mu_r_lookup.weight"][1] - w["mu_r_lookup.weight"][0]
    sn = w["mu_n_lookup.weight"][1] - w["mu_n_lookup.weight"][0]
    _, tok_ref = fo.global_decoder(w, torch.cat([mu_r + 0.5 * sr, mu_n + 0.5 * sn, c], 1), steps)
    assert torch.equal(toks.cpu(), tok_ref)
