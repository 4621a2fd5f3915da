"""Drop-in mirror of the step functions of reference trainer_gmm.py (:100-303).

The reference defines these at module level and closes over the globals `model`, `optimizer`
and `args`; so does this module -- call `configure(model, optimizer, args)` (or assign the
globals) and use `train` / `evaluate` / `loss_function` with the reference's signatures."""
from __future__ import annotations

import torch
from torch.distributions import Normal

from . import _steps
from .ops import ids_to_onehot

model = None
optimizer = None
args = {"beta": 0.2, "lr": 1e-3}


def configure(model_, optimizer_=None, args_=None):
    global model, optimizer, args
    model, optimizer = model_, optimizer_
    if args_ is not None:
        args = args_


def std_normal(shape):
    """trainer_gmm.py:101-106."""
    dev = next(model.parameters()).device if model is not None else "cuda"
    return Normal(torch.zeros(shape, device=dev), torch.ones(shape, device=dev))


def loss_function(out, d, r_out, r, n_out, n, dis, qy_x_out, logLogit_out, step, beta=.1,
                  is_supervised=False, y_label=None):
    """trainer_gmm.py:109-196."""
    return _steps.gm_loss(model, out, d, r_out, r, n_out, n, dis, qy_x_out, logLogit_out, step, beta,
                          is_supervised, y_label)


def latent_regularized_loss_function(z_out, r, n):
    """trainer_gmm.py:199-217."""
    return _steps.latent_reg(z_out, r, n)


def _forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density, is_supervised, y_label):
    res = model(d_oh, r_oh, n_oh, c)
    output, dis, z_out, logLogit_out, qy_x_out, y_out = res
    out, r_out, n_out, _, _ = output
    terms = loss_function(out, d, r_out, r, n_out, n, dis, qy_x_out, logLogit_out, step, beta=args['beta'],
                          is_supervised=is_supervised, y_label=y_label)
    l_r, l_n = latent_regularized_loss_function(z_out, r_density, n_density)
    loss = terms[0] + l_r + l_n
    return loss, terms, l_r, l_n


def _pack(loss, terms, l_r, l_n):
    _, CE_X, CE_R, CE_N, lat_r, lat_n, cls_r, cls_n = terms
    return _steps.to_floats(loss, CE_X, CE_R, CE_N, l_r, l_n, lat_r + lat_n, cls_r + cls_n)


def train(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density, is_supervised=False, y_label=None):
    """trainer_gmm.py:220-258 -> (step+1, (loss, CE_X, CE_R, CE_N, l_r, l_n, kld_latent, kld_class))."""
    optimizer.zero_grad()
    loss, terms, l_r, l_n = _forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density,
                                            is_supervised, y_label)
    _steps.optimise(model, optimizer, loss)
    step += 1
    return step, _pack(loss, terms, l_r, l_n)


def evaluate(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density, is_supervised=False, y_label=None):
    """trainer_gmm.py:261-293 (the reference neither switches to eval mode nor disables grad here)."""
    loss, terms, l_r, l_n = _forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density,
                                            is_supervised, y_label)
    return _pack(loss, terms, l_r, l_n)


def convert_to_one_hot(input, dims):
    """trainer_gmm.py:296-303."""
    return ids_to_onehot(input.cuda(), dims)


# ------------------------------------------------------------------------------------------------
# epoch loop + checkpoints (reference trainer_gmm.py:306-467, :46-50, :458-463)
# ------------------------------------------------------------------------------------------------
EVENT_DIMS, RHYTHM_DIMS, NOTE_DIMS = 342, 3, 16
_TERMS = ("loss", "CE_X", "CE_R", "CE_N", "l_r", "l_n", "kld_latent", "kld_class")


def _run_loader(step, loader, is_supervised, training):
    """One pass over a DataLoader in the reference's batch-tuple layout; returns (step, mean of the 8 terms)."""
    dev = next(model.parameters()).device
    tot = [0.0] * 8
    nb = 0
    for x in loader:
        if is_supervised:                       # VGMIDIDataset tuple (ptb_v2.py:489)
            d, r, n, c, a, v, r_density, n_density = x
        else:                                   # YamahaDataset tuple (ptb_v2.py:436)
            d, r, n, c, r_density, n_density = x
            a = None
        d, r, n, c = d.to(dev).long(), r.to(dev).long(), n.to(dev).long(), c.to(dev).float()
        d_oh, r_oh, n_oh = (convert_to_one_hot(t, k) for t, k in ((d, EVENT_DIMS), (r, RHYTHM_DIMS), (n, NOTE_DIMS)))
        if training:
            step, out = train(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density, is_supervised=is_supervised,
                              y_label=a)
        else:                                   # the reference evaluates with step - 1 (trainer_gmm.py:357,418)
            out = evaluate(step - 1, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density, is_supervised=is_supervised,
                           y_label=a)
        tot = [t + o for t, o in zip(tot, out)]
        nb += 1
    return step, dict(zip(_TERMS, (t / max(nb, 1) for t in tot)))


def training_phase(step, loaders, n_epochs=None, save_path=None, log=print):
    """Epoch loop of reference trainer_gmm.py:306-467.  `loaders` = dict with the reference's four DataLoaders:
    'vgm_train', 'vgm_val' (VGMIDI, supervised pass first) and 'train', 'val' (Yamaha, unsupervised); either pair
    may be absent.  Per epoch: supervised train + val, unsupervised train + val, means of the 8 loss terms,
    then the state_dict is saved (same on-disk format as the reference: a plain fp32 CPU state_dict).
    Returns (step, history)."""
    n_epochs = args.get("n_epochs", 1) if n_epochs is None else n_epochs
    history = []
    for ep in range(1, n_epochs + 1):
        rec = {"epoch": ep}
        for tag, sup in (("vgm", True), ("", False)):
            tr, va = loaders.get(f"{tag}_train" if tag else "train"), loaders.get(f"{tag}_val" if tag else "val")
            if tr is None:
                continue
            step, m_tr = _run_loader(step, tr, sup, True)
            m_va = _run_loader(step, va, sup, False)[1] if va is not None else None
            rec["vgmidi" if sup else "yamaha"] = {"train": m_tr, "val": m_va}
            if log:
                log("epoch {} {}: batch loss {:.5f}  {}".format(ep, "vgmidi" if sup else "yamaha", m_tr["loss"],
                                                               "" if m_va is None else "{:.5f}".format(m_va["loss"])))
        if save_path:
            save_state_dict(model, save_path)
        history.append(rec)
    return step, history


def save_state_dict(model_, path):
    """torch.save(model.cpu().state_dict(), path) of the reference (trainer_gmm.py:458) without moving the model."""
    torch.save({k: v.detach().cpu() for k, v in model_.state_dict().items()}, path)


def save_training_state(path, model_, optimizer_, step):
    """Full resume point (the reference saves the weights only: optimiser moments, step and RNG are lost, so its
    beta annealing restarts -- SURVEY section 5).  Weights stay loadable on their own under 'model'."""
    torch.save({"model": {k: v.detach().cpu() for k, v in model_.state_dict().items()},
                "optimizer": None if optimizer_ is None else {k: (v.detach().cpu() if torch.is_tensor(v) else v)
                                                              for k, v in optimizer_.state_dict().items()},
                "step": int(step), "rng": torch.get_rng_state()}, path)


def load_training_state(path, model_, optimizer_=None):
    """Inverse of save_training_state (also accepts a bare state_dict file written by the reference)."""
    blob = torch.load(path, map_location="cpu")
    if "model" not in blob or not isinstance(blob["model"], dict):
        model_.load_state_dict(blob)
        return 0
    model_.load_state_dict(blob["model"])
    if optimizer_ is not None and blob.get("optimizer"):
        dev = next(model_.parameters()).device
        optimizer_.load_state_dict({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in blob["optimizer"].items()})
    if blob.get("rng") is not None:
        torch.set_rng_state(blob["rng"])
    return int(blob.get("step", 0))
