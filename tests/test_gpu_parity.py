"""GPU parity of the whole path through the reference-facing API (model classes + trainer mirrors ->
C ABI) against (1) the golden vectors produced by the unmodified reference and (2) the CPU oracle on
seeded inputs at shapes the oracle finishes in seconds.  Tolerance: 1e-3 relative fp32 (north_star),
bit-exact for arg-max token ids."""
import numpy as np
import pytest
import torch

from oracle import fader_oracle as fo

pytestmark = pytest.mark.gpu
RTOL = 1e-3


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def close(a, b, rtol=RTOL, atol=1e-5, what=""):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max() if a.size else 0.0
    lim = atol + rtol * np.abs(b).max() if b.size else atol
    assert err <= lim, f"{what}: max abs err {err:.3e} > {lim:.3e}"


# Both arithmetic modes that claim the 1e-3 bar: "f32" (fp32 FMA kernels) and "bf16x3" (the tcgen05 kernels with hi / lo
# bf16 operand planes, three plane products per product).  The tensor-core kernels need H % 64 == 0.
@pytest.fixture(params=["f32", "bf16x3"])
def prec(request):
    return request.param


def build_model(variant, H, Z, K, weights, dev, prec="f32"):
    import fadernets_b200 as fn
    if prec != "f32" and H % 64:
        pytest.skip("tensor-core path needs H % 64 == 0")
    if variant == "gmvae":
        m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    else:
        m = fn.MusicAttrRegVAE(342, 3, 16, 24, H, Z, 32)
    m.load_state_dict(weights)           # strict
    return m.to(dev).train().set_precision(prec)


def fixed_noise(model, *eps):
    it = iter(eps)
    model._draw_eps = lambda B, Z, d: next(it).to(d)
    model.host_rng = False


def golden_batch(g, dev):
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    from fadernets_b200 import trainer_gmm
    d, r, n, c = t("d"), t("r"), t("n"), t("c")
    oh = [trainer_gmm.convert_to_one_hot(x, k) for x, k in ((d, 342), (r, 3), (n, 16))]
    return d, r, n, c, oh


def test_golden_forward_losses_grads(golden, dev, prec):
    from fadernets_b200 import trainer, trainer_gmm
    g = golden
    H, Z, K = int(g["H"]), int(g["Z"]), int(g["K"])
    model = build_model(g["variant"], H, Z, K, g["weights"], dev, prec)
    d, r, n, c, oh = golden_batch(g, dev)
    # (a) seeded host RNG reproduces the reference's noise draw for draw
    torch.manual_seed(g["seed"] + 3)
    model.zero_grad_flat()
    res = model(*oh, c)
    close(res[2][0], g["z_r"], what="z_r (host-RNG parity)")
    close(res[2][1], g["z_n"], what="z_n (host-RNG parity)")
    if g["variant"] == "gmvae":
        trainer_gmm.configure(model, None, {"beta": 0.2})
        (out, r_out, n_out, _, _), dis, z_out, ll, qy, y = res
        terms = trainer_gmm.loss_function(out, d, r_out, r, n_out, n, dis, qy, ll, 20000, beta=0.2)
        names = ("loss", "CE_X", "CE_R", "CE_N", "kld_lat_r", "kld_lat_n", "kld_cls_r", "kld_cls_n")
        for a in "rn":
            i = "rn".index(a)
            close(ll[i], g[f"logLogit_{a}"], rtol=1e-4, atol=1e-2, what="logLogit")
            close(qy[i], g[f"qy_x_{a}"], what="qy_x")
            assert np.array_equal(y[i].cpu().numpy(), g[f"y_{a}"])
        l_r, l_n = trainer_gmm.latent_regularized_loss_function(z_out, g["r_density"], g["n_density"])
    else:
        trainer.configure(model, None, {"beta": 0.2}, step_=20000)
        (out, r_out, n_out), dis, z_out = res
        terms = trainer.loss_function(out, d, r_out, r, n_out, n, dis, beta=0.2)
        names = ("loss", "CE_X", "CE_R", "CE_N")
        l_r, l_n = trainer.latent_regularized_loss_function(z_out, g["r_density"], g["n_density"])
    close(out, g["out"], what="out"); close(r_out, g["r_out"], what="r_out"); close(n_out, g["n_out"], what="n_out")
    for i, a in enumerate("rn"):
        close(dis[i].mean, g[f"mu_{a}"], what="mu"); close(dis[i].stddev, g[f"scale_{a}"], what="scale")
    for nm, tv in zip(names, terms):
        close(tv, g["loss/" + nm], what=nm)
    close(l_r, g["loss/l_r"], what="l_r"); close(l_n, g["loss/l_n"], what="l_n")
    total = terms[0] + l_r + l_n
    close(total, g["loss/total"], what="total")
    total.backward()
    bad = []
    for k in g["live"].tolist():
        ref = g["grad/" + k]
        got = dict(model.named_parameters())[k].grad.cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-6)
        if np.abs(got - ref).max() > RTOL * scale + 1e-7:
            bad.append((k, float(np.abs(got - ref).max() / scale)))
    assert not bad, bad
    for n_, p in model.named_parameters():          # dead / frozen parameters stay gradient-free
        if n_ not in g["live"].tolist():
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n_


def test_golden_loss_branches(golden, dev, prec):
    from fadernets_b200 import trainer_gmm
    g = golden
    if g["variant"] != "gmvae":
        pytest.skip("GM-VAE only")
    model = build_model("gmvae", int(g["H"]), int(g["Z"]), int(g["K"]), g["weights"], dev, prec)
    trainer_gmm.configure(model, None, {"beta": 0.2})
    d, r, n, c, oh = golden_batch(g, dev)
    for tag, st in (("neg_beta", 5000), ("zero_beta", 10)):
        fixed_noise(model, torch.from_numpy(g["eps_r"]), torch.from_numpy(g["eps_n"]))
        (out, r_out, n_out, _, _), dis, z_out, ll, qy, y = model(*oh, c)
        t = trainer_gmm.loss_function(out, d, r_out, r, n_out, n, dis, qy, ll, st, beta=0.2)
        close(t[0], g[f"loss_{tag}/loss"], what=tag)
    # supervised branch incl. gradients
    fixed_noise(model, torch.from_numpy(g["eps_r"]), torch.from_numpy(g["eps_n"]))
    model.zero_grad_flat()
    (out, r_out, n_out, _, _), dis, z_out, ll, qy, y = model(*oh, c)
    yl = torch.from_numpy(g["y_label"]).to(dev)
    t = trainer_gmm.loss_function(out, d, r_out, r, n_out, n, dis, qy, ll, 20000, beta=0.2, is_supervised=True, y_label=yl)
    close(t[0], g["loss_sup/loss"], what="sup loss"); close(t[4], g["loss_sup/kld_lat_r"], what="sup kld r")
    close(t[5], g["loss_sup/kld_lat_n"], what="sup kld n")
    l_r, l_n = trainer_gmm.latent_regularized_loss_function(z_out, g["r_density"], g["n_density"])
    (t[0] + l_r + l_n).backward()
    for k in ("mu_r.weight", "mu_r_lookup.weight", "gru_n.weight_hh_l0_reverse", "grucell_g.weight_ih"):
        ref = g["grad_sup/" + k]
        got = dict(model.named_parameters())[k].grad.cpu().numpy()
        assert np.abs(got - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-6) + 1e-7, k


def test_golden_two_train_steps(golden, dev, prec):
    """Two full train() calls from the reference's initial weights: reported scalars and updated weights."""
    import fadernets_b200 as fn
    from fadernets_b200 import trainer, trainer_gmm
    g = golden
    model = build_model(g["variant"], int(g["H"]), int(g["Z"]), int(g["K"]), g["weights"], dev, prec)
    opt = fn.FusedAdam(model, lr=1e-3)
    d, r, n, c, oh = golden_batch(g, dev)
    torch.manual_seed(g["seed"] + 4)
    outs, step = [], 20000
    for it in range(2):
        if g["variant"] == "gmvae":
            trainer_gmm.configure(model, opt, {"beta": 0.2, "lr": 1e-3})
            step, o = trainer_gmm.train(step, *oh, d, r, n, c, g["r_density"], g["n_density"])
        else:
            trainer.configure(model, opt, {"beta": 0.2, "lr": 1e-3}, step_=step)
            step, o = trainer.train(step, *oh, d, r, n, c, g["r_density"], g["n_density"])
        outs.append(o)
    close(np.array(outs), g["train/outputs"], rtol=2e-3, atol=1e-4, what="train() outputs")
    sd = model.state_dict()
    for k in g["live"].tolist():
        if k in ("linear_out_r.bias", "linear_out_n.bias"):
            # time-axis log-softmax is invariant to a per-class constant: these gradients are
            # mathematically zero and Adam turns their rounding noise into +-lr moves.
            continue
        upd_ref = g["w2/" + k] - g["w/" + k]
        upd = sd[k].cpu().numpy() - g["w/" + k]
        assert np.abs(upd - upd_ref).max() <= 2.5e-4, k            # lr * 2 steps = 2e-3 max move
        assert np.mean(np.abs(upd - upd_ref) > 2e-5) < 0.01, k


def test_golden_greedy_decode(golden, dev, prec):
    g = golden
    model = build_model(g["variant"], int(g["H"]), int(g["Z"]), int(g["K"]), g["weights"], dev, prec)
    model.eval()
    zc = torch.cat([torch.from_numpy(g["z_r"]), torch.from_numpy(g["z_n"]), torch.from_numpy(g["c"])], 1).to(dev)
    steps = g["decode/tokens"].shape[1]
    with torch.no_grad():
        out = model.global_decoder(zc, steps)
    assert np.array_equal(out.argmax(-1).cpu().numpy(), g["decode/tokens"]), "greedy tokens differ"
    close(out, g["decode/logp"], what="decode log-probs")
    lp, toks = model.decode_greedy(zc, steps, return_logp=False)
    assert lp is None and np.array_equal(toks.cpu().numpy(), g["decode/tokens"])
    # encode() / encoder() surface
    oh = torch.nn.functional.one_hot(torch.from_numpy(g["d"]), 342).float().to(dev)
    dis = model.encode(oh) if g["variant"] == "gmvae" else model.encoder(oh)
    close(dis[0].mean, g["mu_r"], what="encode mu_r"); close(dis[1].stddev, g["scale_n"], what="encode scale_n")


@pytest.mark.parametrize("variant,H,Z,K,B,T", [("gmvae", 64, 32, 2, 70, 33), ("vae", 128, 16, 0, 9, 40),
                                               ("gmvae", 256, 128, 2, 16, 24)])
def test_oracle_train_step(dev, prec, variant, H, Z, K, B, T):
    """One train step vs the CPU oracle at mid sizes (ragged batch, pad tail, K3 chunks not multiple of 64)."""
    import fadernets_b200 as fn
    from fadernets_b200 import trainer, trainer_gmm
    w = fo.init_weights(H, Z, variant, max(K, 1), seed=5)
    model = build_model(variant, H, Z, K, w, dev, prec)
    opt = fn.FusedAdam(model, lr=1e-3)
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=6, pad_tail=True)
    g = torch.Generator().manual_seed(8)
    er, en = torch.randn(B, Z, generator=g), torch.randn(B, Z, generator=g)
    scal, grads32, res = fo.loss_and_grads(w, variant, (d, r, n, c, rd, nd), er, en, 20000, 0.2)
    # Gradients are compared with the SAME restatement evaluated in fp64: the fp32 CPU evaluation is itself 7-8e-4 (of the
    # tensor's max) away from it on the worst-conditioned tensors of the H = 256 case (mu_n.weight, the reverse encoders'
    # weight_ih), i.e. it sits at the 1e-3 bar on its own.  Loss values are compared with the fp32 evaluation.
    w64 = {k: v.double() for k, v in w.items()}
    _, grads, _ = fo.loss_and_grads(w64, variant, (d, r, n, c.double(), rd, nd), er.double(), en.double(), 20000, 0.2)
    for k in grads:                       # the two evaluations of the oracle agree to fp32 noise
        assert float((grads32[k].double() - grads[k]).abs().max()) <= 2e-3 * max(float(grads[k].abs().max()), 1e-6) + 1e-7, k
    fixed_noise(model, er, en)
    dd, rr, nn_, cc = d.to(dev), r.to(dev), n.to(dev), c.to(dev)
    opt.zero_grad()
    if variant == "gmvae":
        trainer_gmm.configure(model, opt, {"beta": 0.2})
        loss, terms, l_r, l_n = trainer_gmm._forward_losses(20000, dd, rr, nn_, dd, rr, nn_, cc, rd, nd, False, None)
        close(terms[4], scal["kld_lat_r"], what="kld_lat_r"); close(terms[7], scal["kld_cls_n"], what="kld_cls_n")
    else:
        trainer.configure(model, opt, {"beta": 0.2}, step_=20000)
        loss, *_rest = trainer._forward_losses(dd, rr, nn_, dd, rr, nn_, cc, rd, nd)
    close(loss, scal["loss"], what="loss")
    loss.backward()
    bad = []
    params = dict(model.named_parameters())
    for k, ref in grads.items():
        got = params[k].grad.cpu().double()
        scale = max(float(ref.abs().max()), 1e-6)
        e = float((got - ref).abs().max())
        if e > RTOL * scale + 1e-7:
            bad.append((k, e / scale))
    assert not bad, bad
