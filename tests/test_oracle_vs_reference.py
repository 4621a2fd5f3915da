"""Live check of the oracle against the reference code + its SHIPPED checkpoints
(params/*.pt, H=512 Z=128 K=2).  Runs only where /root/reference exists (the build
container); skipped on the GPU box."""
import os

import numpy as np
import pytest
import torch

from oracle import fader_oracle as fo
from oracle import gen_golden as gg

pytestmark = pytest.mark.skipif(not os.path.isdir(gg.REF), reason="reference checkout not present")


@pytest.mark.parametrize("variant,ckpt", [("gmvae", "music_attr_vae_reg_gmm.pt"),
                                          ("vae", "music_attr_vae_reg_vanilla.pt")])
def test_shipped_checkpoint_forward_loss_decode(variant, ckpt):
    gmm_model, model_v2 = gg.load_reference()
    sd = torch.load(os.path.join(gg.REF, "params", ckpt), map_location="cpu")
    if variant == "gmvae":
        model = gmm_model.MusicAttrRegGMVAE(342, 3, 16, 24, 512, 128, 32, n_component=2)
        trainer = "trainer_gmm.py"
    else:
        model = model_v2.MusicAttrRegVAE(342, 3, 16, 24, 512, 128, 32)
        trainer = "trainer.py"
    model.load_state_dict(sd)                                   # strict
    assert set(sd) == set(fo.param_shapes(512, 128, variant, 2))   # (the shipped gmm file has an older key order)
    model.train()
    ns = gg.extract_step_functions(trainer, gg.trainer_namespace(model, None, dict(beta=0.2), step=20000))
    B, T = 3, 20
    d, r, n, c, rd, nd = fo.synth_batch(B, T, seed=1)
    oh = [ns["convert_to_one_hot"](x, k) for x, k in ((d, 342), (r, 3), (n, 16))]
    torch.manual_seed(0)
    er, en = fo.draw_eps(B, 128, T)
    torch.manual_seed(0)
    with torch.no_grad():
        res = model(*oh, c)
        if variant == "gmvae":
            output, dis, z_out, ll, qy, y = res
            terms = ns["loss_function"](output[0], d, output[1], r, output[2], n, dis, qy, ll, 20000, beta=0.2)
        else:
            output, dis, z_out = res
            terms = ns["loss_function"](output[0], d, output[1], r, output[2], n, dis, beta=0.2)
        lr_, ln_ = ns["latent_regularized_loss_function"](z_out, rd, nd)
    w = {k: v.float() for k, v in sd.items()}
    mine = fo.forward(w, variant, d, r, n, c, er, en)
    np.testing.assert_allclose(mine["out"].numpy(), output[0].numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(mine["z_r"].numpy(), z_out[0].numpy(), rtol=1e-4, atol=1e-5)
    if variant == "gmvae":
        t = fo.loss_gmvae(w, mine, d, r, n, 20000, 0.2)
    else:
        t = fo.loss_vae(mine, d, r, n, 20000, 0.2)
    for a, b in zip(t, terms):
        np.testing.assert_allclose(float(a), float(b.reshape(-1)[0]), rtol=1e-4)
    l = fo.latent_reg(mine["z_r"], mine["z_n"], rd, nd)
    np.testing.assert_allclose([float(l[0]), float(l[1])], [float(lr_), float(ln_)], rtol=1e-4)
    # greedy decode: token ids bit-exact
    model.eval()
    zc = torch.cat([z_out[0], z_out[1], c], 1)
    with torch.no_grad():
        dec = model.global_decoder(zc, steps=24)
    _, toks = fo.global_decoder(w, zc, 24)
    assert np.array_equal(toks.numpy(), dec.argmax(-1).numpy())
