"""The oracle's restatement of the sibling models (model_v2.py:174-586: MusicAttrSingleVAE, MusicAttrCVAE,
MusicAttrFaderNets with their trainers' loss functions) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py make_sibling_case): forward outputs, every loss term, every gradient, greedy tokens."""
import os

import numpy as np
import pytest
import torch

from conftest import SIBLING_GOLDEN_FILES
from oracle import fader_oracle as fo


def load(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["kind"] = os.path.basename(path).split("_")[1]
    g["weights"] = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w/")}
    return g


@pytest.mark.parametrize("path", SIBLING_GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in SIBLING_GOLDEN_FILES])
def test_oracle_siblings_match_reference(path):
    g = load(path)
    kind, w = g["kind"], g["weights"]
    d, c = torch.from_numpy(g["d"]), torch.from_numpy(g["c"])
    eps = torch.from_numpy(g["eps"])
    mr = torch.from_numpy(g["mask_r"]) if "mask_r" in g else None
    mn = torch.from_numpy(g["mask_n"]) if "mask_n" in g else None
    scal, grads, res = fo.sibling_loss_and_grads(w, kind, d, c, g["r_density"], g["n_density"], eps, 20000, 0.2, mr, mn)
    for k in ("out", "mu", "scale", "z"):
        assert np.allclose(res[k].numpy(), g[k], rtol=1e-4, atol=1e-5), k
    if kind == "fader":
        assert np.allclose(res["r_out"].numpy(), g["r_out"], rtol=1e-4, atol=1e-6)
        assert np.allclose(res["n_out"].numpy(), g["n_out"], rtol=1e-4, atol=1e-6)
    for k, v in g.items():
        if k.startswith("loss/"):
            assert abs(float(scal[k[5:]]) - float(v)) <= 1e-4 * max(1.0, abs(float(v))), (k, float(scal[k[5:]]), float(v))
    n = 0
    for k, v in g.items():
        if k.startswith("grad/"):
            assert np.allclose(grads[k[5:]].numpy(), v, rtol=1e-3, atol=1e-6), k
            n += 1
    assert n >= 20
    # eval-mode greedy decode from the same latent: tokens bit-exact
    lp, tok = fo.global_decoder(w, torch.from_numpy(g["z"]), g["decode/tokens"].shape[1])
    assert np.array_equal(tok.numpy(), g["decode/tokens"])
    assert np.allclose(lp.numpy(), g["decode/logp"], rtol=1e-4, atol=1e-4)
