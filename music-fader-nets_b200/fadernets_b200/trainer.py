"""Drop-in mirror of the step functions of reference trainer.py (:80-196, vanilla VAE).

As in the reference, `loss_function` reads the module-global `step` (which the reference never
advances, so beta0 stays 0 there); `configure(..., step=)` lets a caller set it."""
from __future__ import annotations

import torch
from torch.distributions import Normal

from . import _steps
from .ops import ids_to_onehot

model = None
optimizer = None
args = {"beta": 0.2, "lr": 1e-3}
step = 0


def configure(model_, optimizer_=None, args_=None, step_=0):
    global model, optimizer, args, step
    model, optimizer, step = model_, optimizer_, step_
    if args_ is not None:
        args = args_


def std_normal(shape):
    """trainer.py:80-85."""
    dev = next(model.parameters()).device if model is not None else "cuda"
    return Normal(torch.zeros(shape, device=dev), torch.ones(shape, device=dev))


def loss_function(out, d, r_out, r, n_out, n, dis, beta=.1):
    """trainer.py:87-114."""
    return _steps.vae_loss(out, d, r_out, r, n_out, n, dis, step, beta)


def latent_regularized_loss_function(z_out, r, n):
    """trainer.py:117-132."""
    return _steps.latent_reg(z_out, r, n)


def _forward_losses(d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    output, dis, z_out = model(d_oh, r_oh, n_oh, c)
    out, r_out, n_out = output
    loss, CE_X, CE_R, CE_N = loss_function(out, d, r_out, r, n_out, n, dis, beta=args['beta'])
    l_r, l_n = latent_regularized_loss_function(z_out, r_density, n_density)
    return loss + l_r + l_n, CE_X, CE_R, CE_N, l_r, l_n


def train(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    """trainer.py:135-162 -> (step+1, (loss, CE_X, CE_R, CE_N, l_r, l_n)).  NB: like the reference the
    local `step` argument only counts; the beta schedule reads the module global."""
    optimizer.zero_grad()
    terms = _forward_losses(d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density)
    _steps.optimise(model, optimizer, terms[0])
    step += 1
    return step, _steps.to_floats(*terms)


def evaluate(d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    """trainer.py:165-186."""
    return _steps.to_floats(*_forward_losses(d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density))


def convert_to_one_hot(input, dims):
    """trainer.py:189-196."""
    return ids_to_onehot(input.cuda(), dims)
