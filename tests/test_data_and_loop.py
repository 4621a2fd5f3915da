"""the dataset tuple layouts either side of the step (CPU) and the epoch loop +
checkpoint round trip through the accelerated step (GPU)."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader


def test_dataset_tuple_layouts_and_splits():
    from fadernets_b200 import data as D
    y = D.synthetic_yamaha(50, 24, seed=0)
    parts = {m: D.YamahaDataset(*y, mode=m) for m in ("train", "val", "test")}
    assert [len(parts[m]) for m in ("train", "val", "test")] == [40, 5, 5]          # 80 / 10 / 10 (ptb_v2.py:409)
    x, r, n, c, rd, nd = parts["train"][3]
    assert x.shape == (24,) and r.shape == (24,) and n.shape == (24,) and c.shape == (24,)
    assert rd == pytest.approx(float((y[1][3] == 1).mean())) and nd == pytest.approx(float(y[2][3].mean()))
    assert x[-1] == 0 and 1 in x                                                     # EOS then zero padding
    v = D.synthetic_vgmidi(40, 20, seed=1)
    dv = D.VGMIDIDataset(v[0], v[1], v[2], v[5], v[3], v[4], mode="train")
    assert len(dv) == 36                                                            # 90 / 5 / 5 (ptb_v2.py:448)
    x, r, n, c, a, val, rd, nd = dv[0]
    assert x.shape == dv[1][0].shape and x.dtype == torch.float32                   # padded to a common length
    assert a in (0.0, 1.0) and -1 <= val <= 1
    first = v[0][0]
    assert x[len(first) - 1] == 1 and x[len(first)] == first[-1]                    # np.insert(k, -1, 1): EOS before the last token
    batch = next(iter(DataLoader(dv, batch_size=8)))
    assert len(batch) == 8 and batch[0].shape[0] == 8 and batch[6].dtype == torch.float64


@pytest.mark.gpu
def test_training_phase_and_checkpoint_round_trip(lib, tmp_path):
    import fadernets_b200 as fn
    from fadernets_b200 import data as D, trainer_gmm as T
    dev = torch.device("cuda:0")

    def make():
        torch.manual_seed(0)
        m = fn.MusicAttrRegGMVAE(342, 3, 16, 24, 32, 16, 32, n_component=2).to(dev).train()
        return m, fn.FusedAdam(m, lr=1e-3)

    y = D.synthetic_yamaha(40, 12, seed=0)
    v = D.synthetic_vgmidi(40, 12, seed=1)
    loaders = {"train": DataLoader(D.YamahaDataset(*y, mode="train"), batch_size=8),
               "val": DataLoader(D.YamahaDataset(*y, mode="val"), batch_size=8),
               "vgm_train": DataLoader(D.VGMIDIDataset(v[0], v[1], v[2], v[5], v[3], v[4], mode="train"), batch_size=12),
               "vgm_val": DataLoader(D.VGMIDIDataset(v[0], v[1], v[2], v[5], v[3], v[4], mode="val"), batch_size=12)}
    model, opt = make()
    T.configure(model, opt, {"beta": 0.2, "lr": 1e-3, "n_epochs": 2})
    torch.manual_seed(5)
    path = str(tmp_path / "w.pt")
    step, hist = T.training_phase(0, loaders, save_path=path, log=None)
    assert step == 2 * (3 + 4)                      # 36 VGMIDI items / 12 + 32 Yamaha items / 8, two epochs
    for rec in hist:
        for part in ("vgmidi", "yamaha"):
            for split in ("train", "val"):
                assert all(np.isfinite(list(rec[part][split].values()))), rec
    # the weight file has the reference's format: a plain fp32 CPU state_dict with the 80 reference keys
    sd = torch.load(path)
    assert len(sd) == 80 and all(t.device.type == "cpu" and t.dtype == torch.float32 for t in sd.values())

    # full resume point: (weights, Adam moments, step, RNG) -> the continued run matches the uninterrupted one
    ck = str(tmp_path / "state.pt")
    T.save_training_state(ck, model, opt, step)
    step_a, ha = T.training_phase(step, {"train": loaders["train"]}, n_epochs=1, log=None)
    w_a = {k: t.clone() for k, t in model.state_dict().items()}
    model2, opt2 = make()
    T.configure(model2, opt2, {"beta": 0.2, "lr": 1e-3})
    step_b = T.load_training_state(ck, model2, opt2)
    assert step_b == step
    step_b, hb = T.training_phase(step_b, {"train": loaders["train"]}, n_epochs=1, log=None)
    assert step_a == step_b and ha[0]["yamaha"]["train"] == hb[0]["yamaha"]["train"]
    for k, t in model2.state_dict().items():
        assert torch.equal(t, w_a[k]), k


@pytest.mark.gpu
def test_training_phase_against_oracle(lib):
    """The epoch loop (reference trainer_gmm.py:306-467: supervised VGMIDI pass, then the unsupervised Yamaha pass,
    evaluate() with step - 1) against the CPU oracle driven batch by batch with the same CPU-generator draws: the
    per-epoch means of the 8 terms and the final weights."""
    import fadernets_b200 as fn
    from fadernets_b200 import data as D, trainer_gmm as T
    from oracle import fader_oracle as fo
    dev = torch.device("cuda:0")
    H, Z, K = 16, 8, 2
    w = fo.init_weights(H, Z, "gmvae", K, seed=1)
    model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    model.load_state_dict(w)
    model = model.to(dev).train()
    opt = fn.FusedAdam(model, lr=1e-3)
    y = D.synthetic_yamaha(20, 10, seed=0)
    v = D.synthetic_vgmidi(20, 10, seed=1)
    loaders = {"train": DataLoader(D.YamahaDataset(*y, mode="train"), batch_size=8),
               "val": DataLoader(D.YamahaDataset(*y, mode="val"), batch_size=8),
               "vgm_train": DataLoader(D.VGMIDIDataset(v[0], v[1], v[2], v[5], v[3], v[4], mode="train"), batch_size=6),
               "vgm_val": DataLoader(D.VGMIDIDataset(v[0], v[1], v[2], v[5], v[3], v[4], mode="val"), batch_size=6)}
    T.configure(model, opt, {"beta": 0.2, "lr": 1e-3, "n_epochs": 1})
    step0 = 20000
    torch.manual_seed(9)
    step, hist = T.training_phase(step0, loaders, log=None)

    # ---- oracle: same order of batches, same draws (every forward consumes eps_r, eps_n and T coin flips)
    torch.manual_seed(9)
    wo = {k: t.clone() for k, t in w.items()}
    st = fo.AdamState(wo)
    names = ("loss", "CE_X", "CE_R", "CE_N", "l_r", "l_n", "kld_latent", "kld_class")
    ostep = step0
    expect = {}
    for tag, sup in (("vgm", True), ("", False)):
        for split, training in (("train", True), ("val", False)):
            tot, nb = np.zeros(8), 0
            for x in loaders[f"{tag}_{split}" if tag else split]:
                if sup:
                    d, r, n, c, a, val, rd, nd = x
                else:
                    d, r, n, c, rd, nd = x
                    a = None
                d, r, n, c = d.long(), r.long(), n.long(), c.float()
                er, en = fo.draw_eps(d.shape[0], Z, d.shape[1])
                batch = (d, r, n, c, rd.numpy(), nd.numpy())
                ylab = None if a is None else a.long()
                if training:
                    s, _ = fo.train_step(wo, st, "gmvae", batch, er, en, ostep, 0.2, 1e-3, y_label=ylab)
                    ostep += 1
                else:
                    s, _, _ = fo.loss_and_grads(wo, "gmvae", batch, er, en, ostep - 1, 0.2, y_label=ylab)
                row = [float(s["loss"]),
                       float(s["CE_X"]), float(s["CE_R"]), float(s["CE_N"]), float(s["l_r"]), float(s["l_n"]),
                       float(s["kld_lat_r"]) + float(s["kld_lat_n"]), float(s["kld_cls_r"]) + float(s["kld_cls_n"])]
                tot += np.array(row); nb += 1
            expect[("vgmidi" if sup else "yamaha", split)] = dict(zip(names, tot / nb))
    assert step == ostep
    for (part, split), ref in expect.items():
        got = hist[0][part][split]
        for k in names[1:]:
            assert abs(got[k] - ref[k]) <= 1e-3 * max(1.0, abs(ref[k])), (part, split, k, got[k], ref[k])
    sd = model.state_dict()
    for k in ("grucell_g.weight_hh", "gru_r.weight_ih_l0", "mu_n_lookup.weight", "linear_out_g.bias"):
        assert torch.allclose(sd[k].cpu(), wo[k], rtol=1e-3, atol=3e-5), k
