"""GLSR trainer (reference trainer_glsr.py:82-258 on MusicAttrRegVAE): the oracle's restatement of the
regulariser -- four extra teacher-forced 100-step decodes, the attribute approximations with the reference's semantics as
written -- against golden vectors of the UNMODIFIED reference (oracle/gen_golden.py make_glsr_case)."""
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR
from oracle import fader_oracle as fo

PATH = os.path.join(GOLDEN_DIR, "glsr_vae_H16_Z8_B3_T104.npz")
SEED = 81


def load():
    z = np.load(PATH, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["weights"] = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w/")}
    return g


def batch(g):
    return tuple(torch.from_numpy(g[k]) for k in "drnc") + (g["r_density"], g["n_density"])


def test_glsr_losses_and_gradients_match_reference():
    g = load()
    B, T, Z = int(g["B"]), int(g["T"]), int(g["Z"])
    torch.manual_seed(SEED + 3)
    er, en = fo.draw_eps(B, Z, T)
    dr, dn = fo.draw_glsr_deltas(B)
    scal, grads = fo.glsr_loss_and_grads(g["weights"], batch(g), er, en, dr, dn, 20000, 0.2)
    for k in ("CE_X", "CE_R", "CE_N", "l_r", "l_n"):
        np.testing.assert_allclose(float(scal[k]), float(g["loss/" + k]), rtol=2e-5, err_msg=k)
    np.testing.assert_allclose(float(scal["loss"]), float(g["loss/total"]), rtol=2e-5)
    assert float(g["loss/l_n"]) > 0.93                      # the note-density branch is live (0.9189 = a zero finite difference)
    for k in g["live"].tolist():
        if k in ("linear_out_r.bias", "linear_out_n.bias"):  # mathematically zero (time-axis soft-max)
            continue
        ref = g["grad/" + k]
        assert np.abs(grads[k].numpy() - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k


def test_glsr_two_train_steps_replay_reference_rng():
    g = load()
    B, T, Z = int(g["B"]), int(g["T"]), int(g["Z"])
    w = {k: v.clone() for k, v in g["weights"].items()}
    st = fo.AdamState(w)
    torch.manual_seed(SEED + 4)
    for it in range(2):
        er, en = fo.draw_eps(B, Z, T)
        dr, dn = fo.draw_glsr_deltas(B)
        scal, grads = fo.glsr_loss_and_grads(w, batch(g), er, en, dr, dn, 20000 + it, 0.2)
        fo.clip_and_adam(w, grads, st, 1e-3)
        got = [float(scal[k]) for k in ("loss", "CE_X", "CE_R", "CE_N", "l_r", "l_n")]
        np.testing.assert_allclose(got, g["train/outputs"][it], rtol=1e-4, err_msg=f"step {it}")
    assert g["train/outputs"][0][4] > 0.93                  # a separator count changed between z+ and z-: the rhythm branch is live
    for k in g["live"].tolist():
        if k in ("linear_out_r.bias", "linear_out_n.bias"):
            continue
        upd_ref, upd = g["w2/" + k] - g["w/" + k], w[k].numpy() - g["w/" + k]
        # Adam moves a weight by ~lr whatever its gradient's size: a ~zero gradient whose sign flips moves it the other way
        # (<= 2 lr per step); the regulariser's finite differences (/ 2 delta ~ / 0.03) amplify fp32 noise into such flips
        assert np.abs(upd - upd_ref).max() <= 4e-4, k
        assert np.mean(np.abs(upd - upd_ref) > 2e-5) < 0.01, k


def test_glsr_regulariser_is_off_in_the_first_steps():
    g = load()
    B, T, Z = int(g["B"]), int(g["T"]), int(g["Z"])
    torch.manual_seed(SEED + 5)
    er, en = fo.draw_eps(B, Z, T)
    scal, _ = fo.glsr_loss_and_grads(g["weights"], batch(g), er, en, None, None, 10, 0.2)
    ref = g["eval_early/outputs"]
    assert ref[4] == 0.0 and ref[5] == 0.0 and float(scal["l_r"]) == 0.0
    np.testing.assert_allclose([float(scal[k]) for k in ("loss", "CE_X", "CE_R", "CE_N")], ref[:4], rtol=2e-5)


def test_product_rhythm_density_scan_equals_the_loop_restatement():
    """The mirror's vectorised approx_rhythm_density (fadernets_b200/trainer_glsr.py: run sums through cumsum / cummax, pure
    tensor operations, device-agnostic) against the oracle's step-by-step restatement of the reference loop, values AND
    gradients, on synthetic log-probabilities whose time-shift mass straddles the 0.9 threshold (runs of every length,
    separators at step 0, trailing unflushed mass, small masses below 1e-2 that flush as themselves)."""
    from fadernets_b200 import trainer_glsr as tg
    gen = torch.Generator().manual_seed(3)
    for case in range(6):
        B, S, V = 5, 100, 342
        logits = torch.randn(B, S, V, generator=gen, dtype=torch.float64) * 0.3
        boost = torch.where(torch.rand(B, S, generator=gen) < (0.2 + 0.12 * case), 6.5, 2.0).double()
        logits[..., 180:278] += boost.unsqueeze(-1)                       # time-shift mass ~0.99 or ~0.75
        if case % 2:
            logits[0, :, 2:90] -= 6.0                                     # sequence 0's note mass small: runs below 1e-2
        logits[:, 0, 180:278] += 3.0 * (case % 3 == 0)                    # separator at step 0
        lp = torch.log_softmax(logits, -1).requires_grad_(True)
        a = tg.approx_rhythm_density(lp)
        b = fo.glsr_rhythm_density(lp)
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-15), (case, a, b)
        assert float(a.abs().sum()) > 0
        wgt = torch.randn(B, generator=gen, dtype=torch.float64)
        ga, = torch.autograd.grad((a * wgt).sum(), lp, retain_graph=True)
        gb, = torch.autograd.grad((b * wgt).sum(), lp)
        assert torch.allclose(ga, gb, rtol=1e-9, atol=1e-14), case
        assert torch.allclose(tg.approx_note_density(lp), fo.glsr_note_density(lp), rtol=1e-12)


def test_golden_case_is_not_borderline():
    """The separator decisions (time-shift mass >= 0.9) of every decode the golden case runs are at least 3e-5 away from
    the threshold, so fp32 re-association on the GPU cannot flip one."""
    g = load()
    B, T, Z = int(g["B"]), int(g["T"]), int(g["Z"])
    w, (d, r, n, c, _, _) = g["weights"], batch(g)
    torch.manual_seed(SEED + 3)
    er, en = fo.draw_eps(B, Z, T)
    dr, dn = fo.draw_glsr_deltas(B)
    res = fo.forward(w, "vae", d, r, n, c, er, en, True)
    margin = 1.0
    for sign in (1, -1):
        z = res["z_r"].clone(); z[:, 0] += sign * dr
        lp = fo.global_decoder(w, torch.cat([z, res["z_n"], c], 1), fo.GLSR_STEPS, d[:, :fo.GLSR_STEPS])[0]
        mass = torch.softmax(lp, -1)[..., 180:278].sum(-1)
        margin = min(margin, float((mass - 0.9).abs().min()))
    assert margin > 3e-5, margin
