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
