"""Drop-in mirror of the step functions of reference trainer_glsr.py (:82-279): the vanilla `MusicAttrRegVAE` trained with the
GLSR regulariser of Hadjeres et al. as the reference implements it.

The heavy part of the regulariser -- FOUR extra teacher-forced 100-step decodes of the global decoder per step, and their
backward passes -- runs on the same CUDA kernels as everything else (`model.global_decoder` in training mode).  The attribute
approximations on top of the decoded log-probabilities (a few thousand scalars per batch; Python loops with `.item()` control
flow in the reference, :139-171) are a handful of batched device operations here, with the reference's semantics AS WRITTEN:
the running note mass always reads sequence 0 (`played_notes[0][i]`, :154), a flushed mass above 1e-2 counts as `cur / cur`
(value 1, zero gradient), the tail after the last separator is dropped, a density of exactly 0 becomes a constant."""
from __future__ import annotations

import math

import torch
from torch.distributions import Normal

from . import _steps
from .ops import ids_to_onehot

model = None
optimizer = None
args = {"beta": 0.2, "lr": 1e-3}

GLSR_STEPS = 100                 # trainer_glsr.py:187,189,212,214
EPSILON = 1e-2                   # :176
NOTE_ON = (2, 90)                # :124-126  tokens 2..89
TIME_SHIFT = (180, 278)          # :132-134  tokens 180..277


def configure(model_, optimizer_=None, args_=None):
    global model, optimizer, args
    model, optimizer = model_, optimizer_
    if args_ is not None:
        args = args_


def std_normal(shape):
    """trainer_glsr.py:82-87."""
    dev = next(model.parameters()).device if model is not None else "cuda"
    return Normal(torch.zeros(shape, device=dev), torch.ones(shape, device=dev))


def loss_function(out, d, r_out, r, n_out, n, dis, step, beta=.1):
    """trainer_glsr.py:90-115 -> (loss, CE_X, CE_R, CE_N): 5 CE_X + CE_R + CE_N + beta0 KL(q || N(0,1))."""
    return _steps.vae_loss(out, d, r_out, r, n_out, n, dis, step, beta)


def _masses(logp):
    """softmax of the decoder's log-probabilities (the reference re-normalises them, :127,135) summed over the note-on and
    the time-shift tokens: (B,S) each."""
    p = torch.softmax(logp, -1)
    return p[..., NOTE_ON[0]:NOTE_ON[1]].sum(-1), p[..., TIME_SHIFT[0]:TIME_SHIFT[1]].sum(-1)


def approx_note_density(logp):
    """trainer_glsr.py:139-141 -> (B,)."""
    return _masses(logp)[0].sum(1)


def approx_rhythm_density(logp):
    """trainer_glsr.py:143-171 -> (B,), without the per-element host loop.  A maximal run of steps whose time-shift mass is below
    0.9 accumulates the note mass of SEQUENCE 0; the separator that ends the run flushes it -- as 1 if it exceeds 1e-2, else as
    itself (a run whose sum is exactly 0 is not flushed); the density is the flushed total over the summed time-shift mass."""
    notes, shift = _masses(logp)
    sep = shift.detach() >= 0.9                                                   # (B,S)
    acc = torch.cumsum(notes[0].unsqueeze(0) * (~sep).to(notes.dtype), 1)         # note mass of sequence 0 over the non-separator steps
    prev_sep = torch.nn.functional.pad(sep[:, :-1], (1, 0), value=True)           # a separator at step 0 has nothing to flush
    ends_run = sep & ~prev_sep
    # mass accumulated since the previous flush = acc - (acc at the previous run end); acc is non-decreasing
    at_end = torch.where(ends_run, acc, torch.zeros_like(acc))
    before = torch.nn.functional.pad(torch.cummax(at_end, 1).values[:, :-1], (1, 0))
    run = acc - before
    flushed = ends_run & (run.detach() != 0)
    piece = torch.where(run.detach() > 1e-2, torch.ones_like(run), run)           # cur / cur: value 1, gradient 0
    total = torch.where(flushed, piece, torch.zeros_like(run)).sum(1)
    dens = total / shift.sum(1)
    return torch.where(total.detach() != 0, dens, torch.zeros_like(dens))


def latent_regularized_loss_function(z_out, r, n, c):
    """trainer_glsr.py:118-229 -> (l_r, l_n).  `r`, `n` (the batch's attribute values) are unused, as in the reference.
    Draws the finite-difference steps on the CPU default generator like the reference when model.host_rng (so a seeded run
    replays it draw for draw: deltas, then the coin flips of the two decodes, per latent)."""
    z_r, z_n = z_out
    dev = z_r.device
    c = c.to(dev).float()

    def deltas(B):
        u = torch.rand(B) if getattr(model, "host_rng", True) else torch.rand(B, device=dev)
        return ((1 + u) * EPSILON).to(dev)

    def shifted(z, dlt):
        e0 = torch.zeros_like(z)
        e0[:, 0] = dlt
        return z + e0

    def finite_difference(attr, plus, minus, dlt):
        a_p = attr(model.global_decoder(torch.cat(plus + [c], dim=1), steps=GLSR_STEPS))
        a_m = attr(model.global_decoder(torch.cat(minus + [c], dim=1), steps=GLSR_STEPS))
        g = (a_p - a_m) / (2 * dlt)
        return (0.5 * g * g + 0.5 * math.log(2 * math.pi)).mean()            # -log N(g; 0, 1)

    d_r = deltas(z_r.size(0))
    l_r = finite_difference(approx_rhythm_density, [shifted(z_r, d_r), z_n], [shifted(z_r, -d_r), z_n], d_r)
    d_n = deltas(z_n.size(0))
    l_n = finite_difference(approx_note_density, [z_r, shifted(z_n, d_n)], [z_r, shifted(z_n, -d_n)], d_n)
    return l_r, l_n


def _forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    output, dis, z_out = model(d_oh, r_oh, n_oh, c)
    out, r_out, n_out = output
    loss, CE_X, CE_R, CE_N = loss_function(out, d, r_out, r, n_out, n, dis, step, beta=args['beta'])
    zero = torch.zeros((), device=out.device)
    l_r, l_n = zero, zero
    if step > 20:                                     # "apply GLSR after 20 steps of training" (:248-252)
        l_r, l_n = latent_regularized_loss_function(z_out, r_density, n_density, c)
        loss = loss + l_r + l_n
    return loss, CE_X, CE_R, CE_N, l_r, l_n


def train(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    """trainer_glsr.py:232-258 -> (step+1, (loss, CE_X, CE_R, CE_N, l_r, l_n))."""
    optimizer.zero_grad()
    terms = _forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density)
    _steps.optimise(model, optimizer, terms[0])
    step += 1
    return step, _steps.to_floats(*terms)


def evaluate(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density):
    """trainer_glsr.py:261-284."""
    return _steps.to_floats(*_forward_losses(step, d_oh, r_oh, n_oh, d, r, n, c, r_density, n_density))


def convert_to_one_hot(input, dims):
    """trainer_glsr.py:287-294."""
    return ids_to_onehot(input.cuda(), dims)
