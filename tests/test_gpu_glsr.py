"""GLSR trainer mirror (fadernets_b200.trainer_glsr; reference trainer_glsr.py:82-279 on MusicAttrRegVAE) against golden
vectors of the UNMODIFIED reference: loss terms, the regulariser (four extra teacher-forced 100-step decodes through the
CUDA kernels), every gradient, two train() calls incl. updated weights, and the step <= 20 branch -- with the reference's
CPU-generator draws replayed from the seed alone (model noise, decoder coin flips, finite-difference steps)."""
import numpy as np
import pytest
import torch

from test_oracle_glsr import SEED, load

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _setup(g, dev):
    import fadernets_b200 as fn
    from fadernets_b200 import trainer_glsr as T
    m = fn.MusicAttrRegVAE(342, 3, 16, 24, int(g["H"]), int(g["Z"]), 32)
    m.load_state_dict(g["weights"])
    m = m.to(dev).train()
    opt = fn.FusedAdam(m, lr=1e-3)
    T.configure(m, opt, {"beta": 0.2, "lr": 1e-3})
    d, r, n, c = (torch.from_numpy(g[k]).to(dev) for k in "drnc")
    oh = [T.convert_to_one_hot(x, k) for x, k in ((d, 342), (r, 3), (n, 16))]
    return m, opt, T, (d, r, n, c), oh


def test_glsr_step_matches_reference(lib):
    g = load()
    dev = torch.device("cuda:0")
    m, opt, T, (d, r, n, c), oh = _setup(g, dev)
    opt.zero_grad()
    torch.manual_seed(SEED + 3)
    loss, CE_X, CE_R, CE_N, l_r, l_n = T._forward_losses(20000, *oh, d, r, n, c, g["r_density"], g["n_density"])
    for nm, t in (("CE_X", CE_X), ("CE_R", CE_R), ("CE_N", CE_N), ("l_r", l_r), ("l_n", l_n), ("total", loss)):
        e = float(g["loss/" + nm])
        assert abs(float(t) - e) <= RTOL * max(1.0, abs(e)), (nm, float(t), e)
    loss.backward()
    params = dict(m.named_parameters())
    bad = []
    for k in g["live"].tolist():
        ref = g["grad/" + k]
        got = params[k].grad.cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-6)
        if np.abs(got - ref).max() > RTOL * scale + 1e-7:
            bad.append((k, float(np.abs(got - ref).max() / scale)))
    assert not bad, bad


def test_glsr_two_train_steps_and_early_steps(lib):
    g = load()
    dev = torch.device("cuda:0")
    m, opt, T, (d, r, n, c), oh = _setup(g, dev)
    torch.manual_seed(SEED + 4)
    step = 20000
    for it in range(2):
        step, o = T.train(step, *oh, d, r, n, c, g["r_density"], g["n_density"])
        for a, b in zip(o, g["train/outputs"][it]):
            assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (it, o, g["train/outputs"][it])
    assert step == 20002
    sd = m.state_dict()
    for k in g["live"].tolist():
        if k in ("linear_out_r.bias", "linear_out_n.bias"):            # mathematically zero gradients (time-axis soft-max)
            continue
        upd_ref, upd = g["w2/" + k] - g["w/" + k], sd[k].cpu().numpy() - g["w/" + k]
        assert np.abs(upd - upd_ref).max() <= 4e-4, k                  # see tests/test_oracle_glsr.py on Adam sign flips
        assert np.mean(np.abs(upd - upd_ref) > 2e-5) < 0.01, k
    # the regulariser is off for step <= 20 (trainer_glsr.py:250-252)
    m.load_state_dict(g["weights"])
    torch.manual_seed(SEED + 5)
    o = T.evaluate(10, *oh, d, r, n, c, g["r_density"], g["n_density"])
    assert o[4] == 0.0 and o[5] == 0.0
    np.testing.assert_allclose(o[:4], g["eval_early/outputs"][:4], rtol=RTOL)
