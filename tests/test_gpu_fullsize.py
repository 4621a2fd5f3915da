"""Full-size runs (BASELINE configs 2 and 3) checked through size-independent properties -- the CPU oracle cannot run
these shapes in seconds: normalisation of every returned log-probability tensor, run-to-run determinism, independence
of sequences (a half batch reproduces its rows), agreement of the two arithmetic modes, dead parameters untouched;
plus the error behaviour at the limits of the tensor-core path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    return torch.device("cuda:0")


def _batch(B, T, seed, dev):
    g = torch.Generator().manual_seed(seed)
    d = torch.randint(2, 342, (B, T), generator=g)
    d[:, -T // 8:] = 0
    d[:, -T // 8] = 1
    r = torch.randint(0, 3, (B, T), generator=g)
    n = torch.randint(0, 16, (B, T), generator=g)
    c = torch.rand(B, 24, generator=g)
    rd, nd = (r == 1).double().mean(1), n.double().mean(1)
    return [t.to(dev) for t in (d, r, n, c)] + [rd, nd]


def _step(model, batch, eps):
    from fadernets_b200 import trainer_gmm
    d, r, n, c, rd, nd = batch
    it = iter(eps)
    model._draw_eps = lambda B_, Z_, d_: next(it)
    model.host_rng = False
    model.zero_grad_flat()
    trainer_gmm.configure(model, None, {"beta": 0.2})
    res = model(d, r, n, c)                                   # ids are accepted in place of the one-hots
    (out, r_out, n_out, _, _), dis, z_out, ll, qy, y = res
    terms = trainer_gmm.loss_function(out, d, r_out, r, n_out, n, dis, qy, ll, 20000, beta=0.2)
    l_r, l_n = trainer_gmm.latent_regularized_loss_function(z_out, rd, nd)
    loss = terms[0] + l_r + l_n
    loss.backward()
    return loss.detach(), out.detach(), r_out.detach(), n_out.detach()


def test_config3_full_size_properties(dev):
    import fadernets_b200 as fn
    B, T, H, Z = 256, 512, 1024, 128
    torch.manual_seed(0)
    model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=2).to(dev).train().set_precision("bf16")
    batch = _batch(B, T, 1, dev)
    g = torch.Generator().manual_seed(2)
    eps = [torch.randn(B, Z, generator=g).to(dev) for _ in range(2)]
    loss1, out, r_out, n_out = _step(model, batch, eps)
    grad1 = model._flat_grad.clone()
    assert torch.isfinite(loss1) and torch.isfinite(grad1).all()
    # every returned tensor is a log-probability over its own axis (vocabulary; TIME for the sub-decoders)
    assert torch.allclose(out[::37].exp().sum(-1), torch.ones_like(out[::37, :, 0]), atol=2e-4)
    assert torch.allclose(r_out.exp().sum(1), torch.ones(B, 3, device=dev), atol=2e-3)
    assert torch.allclose(n_out.exp().sum(1), torch.ones(B, 16, device=dev), atol=2e-3)
    # dead / frozen parameters of the reference stay gradient-free
    for name, p in model.named_parameters():
        if name.startswith(("gru_c.", "gru_d_c.", "c_r.", "c_n.", "mu_c.", "var_c.", "linear_init_c.", "linear_out_c.", "logvar_")):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
    # run-to-run determinism (no float atomics anywhere, fixed-order split-K)
    loss2, out2, _, _ = _step(model, batch, eps)
    gd = (grad1 - model._flat_grad).abs()
    assert torch.equal(loss1, loss2) and torch.equal(out, out2) and torch.equal(grad1, model._flat_grad), \
        ("not run-to-run deterministic", float((loss1 - loss2).abs()), float((out - out2).abs().max()), float(gd.max()),
         int((gd > 0).sum()), int(gd.argmax()))
    # sequences are independent: the first 128 sequences alone reproduce their rows
    half = [t[:128] for t in batch]
    _, out_h, r_h, _ = _step(model, half, [e[:128] for e in eps])
    assert float((out_h - out[:128]).abs().max()) < 2e-2
    assert float((r_h - r_out[:128]).abs().max()) < 2e-2


def test_config2_full_size_modes_agree(dev):
    import fadernets_b200 as fn
    B, T, H, Z = 64, 256, 512, 128
    torch.manual_seed(0)
    model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=2).to(dev).train()
    batch = _batch(B, T, 3, dev)
    g = torch.Generator().manual_seed(4)
    eps = [torch.randn(B, Z, generator=g).to(dev) for _ in range(2)]
    loss32, out32, r32, _ = _step(model, batch, eps)
    g32 = model._flat_grad.clone()
    model.set_precision("bf16")
    loss16, out16, r16, _ = _step(model, batch, eps)
    g16 = model._flat_grad
    assert abs(float(loss16) - float(loss32)) <= 2e-2 * abs(float(loss32))
    assert float((out16 - out32).abs().max()) < 0.15
    cos = float((g16 * g32).sum() / (g16.norm() * g32.norm()))
    assert cos > 0.995, cos


def test_tensor_core_path_limits_fail_loudly(dev):
    import fadernets_b200 as fn
    from fadernets_b200 import FaderNetsError
    torch.manual_seed(0)
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, 96, 16, 32, n_component=2).to(dev).train().set_precision("bf16")
    b = _batch(4, 8, 0, dev)
    with pytest.raises(FaderNetsError, match="64"):            # H % 64 != 0: no silent fallback to another path
        m(b[0], b[1], b[2], b[3])
    m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, 64, 16, 32, n_component=2).to(dev).train().set_precision("bf16")
    # a batch above the 256 rows one chain holds is cut into groups of <= 256 sequences (ops_bf16.GruGroupBf16):
    # same results as the fp32 path within bf16 tolerance, gradients flow to every parameter
    b = _batch(300, 8, 0, dev)
    torch.manual_seed(5)
    res16 = m(b[0], b[1], b[2], b[3])
    out16 = res16[0][0]
    assert out16.shape[0] == 300 and torch.isfinite(out16).all()
    out16.float().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for _, p in m.live_parameters())
    assert float(dict(m.named_parameters())["gru_r.weight_hh_l0"].grad.abs().sum()) > 0
    m.set_precision("f32")
    torch.manual_seed(5)
    out32 = m(b[0], b[1], b[2], b[3])[0][0]
    assert float((out16 - out32).abs().max()) < 5e-2
    # the same batch in bf16x3 mode (chain groups of split tensors): fp32-level agreement, values and gradients
    m.zero_grad_flat()
    torch.manual_seed(5)
    out32 = m(b[0], b[1], b[2], b[3])[0][0]
    out32.sum().backward()
    g32 = {k: p.grad.detach().clone() for k, p in m.live_parameters()}
    m.set_precision("bf16x3")
    m.zero_grad_flat()
    torch.manual_seed(5)
    out3 = m(b[0], b[1], b[2], b[3])[0][0]
    assert float((out3 - out32).abs().max()) <= 1e-3 * float(out32.abs().max())
    out3.sum().backward()
    for k, p in m.live_parameters():
        assert float((p.grad - g32[k]).abs().max()) <= 1e-3 * max(float(g32[k].abs().max()), 1e-6) + 1e-6, k
    m.set_precision("bf16")
    with pytest.raises(FaderNetsError):
        m.set_precision("fp8")
