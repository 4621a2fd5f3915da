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
