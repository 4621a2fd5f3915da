"""Step functions of the sibling trainers (reference trainer_singlevae.py:86-170, trainer_cvae.py:84-135,
trainer_fader.py:84-152) on the accelerated models: same names, signatures and return tuples, grouped per model in a
small namespace class because the three reference scripts reuse the same function names."""
from __future__ import annotations

import torch

from . import _steps
from .ops import LatentRegFn, MseMeanFn, NllMeanFn, StdKlFn


class _Trainer:
    def __init__(self):
        self.model = self.optimizer = None
        self.args = {"beta": 0.1, "lr": 1e-3}

    def configure(self, model, optimizer=None, args=None):
        self.model, self.optimizer = model, optimizer
        if args is not None:
            self.args = args

    def convert_to_one_hot(self, input, dims):
        from . import ops
        return ops.ids_to_onehot(input.cuda(), dims)

    def _base_loss(self, out, d, dis, step, beta, ce_weight, annealed):
        beta0 = _steps.beta_anneal(step, beta)
        CE_X = NllMeanFn.apply(out, d.to(out.device))
        KLD = StdKlFn.apply(dis.mean, dis.stddev)
        return ce_weight * CE_X + (beta0 if annealed else beta) * KLD, CE_X


class SingleVAETrainer(_Trainer):
    """trainer_singlevae.py: loss = 5 CE_X + beta KLD (NOT the annealed beta0, :109) + l_r + l_n on z[:,0], z[:,1]."""

    def loss_function(self, out, d, dis, step, beta=.1):
        return self._base_loss(out, d, dis, step, beta, 5, annealed=False)

    def latent_regularized_loss_function(self, z_out, r, n):
        dev = z_out.device
        return (LatentRegFn.apply(z_out, _steps._attr(r, dev)),
                LatentRegFn.apply(z_out[:, 1:].contiguous(), _steps._attr(n, dev)))

    def _losses(self, step, d_oh, d, c, r_density, n_density):
        out, dis, z = self.model(d_oh, c)
        loss, CE_X = self.loss_function(out, d, dis, step, beta=self.args["beta"])
        l_r, l_n = self.latent_regularized_loss_function(z, r_density, n_density)
        return loss + l_r + l_n, CE_X, l_r, l_n

    def train(self, step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
        self.optimizer.zero_grad()
        terms = self._losses(step, d_oh, d, c, r_density, n_density)
        _steps.optimise(self.model, self.optimizer, terms[0])
        return step + 1, _steps.to_floats(*terms)

    def evaluate(self, step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
        return _steps.to_floats(*self._losses(step, d_oh, d, c, r_density, n_density))


class CVAETrainer(_Trainer):
    """trainer_cvae.py: loss = CE_X + beta0 KLD; r_density / n_density are (B,1) tensors (:122-125)."""

    def loss_function(self, out, d, dis, step, beta=.1):
        return self._base_loss(out, d, dis, step, beta, 1, annealed=True)

    def _losses(self, step, d_oh, r_oh, n_oh, d, c, r_density, n_density):
        out, dis, z = self.model(d_oh, r_oh, n_oh, c, r_density, n_density)
        return self.loss_function(out, d, dis, step, beta=self.args["beta"])

    def train(self, step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
        self.optimizer.zero_grad()
        terms = self._losses(step, d_oh, r_oh, n_oh, d, c, r_density, n_density)
        _steps.optimise(self.model, self.optimizer, terms[0])
        return step + 1, _steps.to_floats(*terms)


class FaderTrainer(CVAETrainer):
    """trainer_fader.py: + lmbda (MSE(r_out, r_density) + MSE(n_out, n_density)), lmbda = min(step / 2000 * 1e-4, 1e-4)."""

    def adversarial_loss(self, step, r_out, n_out, r_density, n_density):
        lmbda = min(step / 2000 * 1e-4, 1e-4)
        dev = r_out.device
        return (lmbda * MseMeanFn.apply(r_out, r_density.to(dev).float()), lmbda * MseMeanFn.apply(n_out, n_density.to(dev).float()))

    def _losses(self, step, d_oh, r_oh, n_oh, d, c, r_density, n_density):
        (out, r_out, n_out), dis, z = self.model(d_oh, r_oh, n_oh, c, r_density, n_density)
        loss, CE_X = self.loss_function(out, d, dis, step, beta=self.args["beta"])
        l_adv_r, l_adv_n = self.adversarial_loss(step, r_out, n_out, r_density, n_density)
        return loss + l_adv_r + l_adv_n, CE_X, l_adv_r, l_adv_n

    def evaluate(self, step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
        return _steps.to_floats(*self._losses(step, d_oh, r_oh, n_oh, d, c, r_density, n_density))
