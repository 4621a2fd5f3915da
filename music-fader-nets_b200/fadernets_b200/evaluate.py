"""The decode drivers around `global_decoder` (reference test_class.py:34-56, 233-254, 282-303 and
arousal_transfer.ipynb cells 11/15/17), batched: the reference shifts and decodes ONE sequence at a time; here
every helper takes a batch so the 8 `shift()` decodes per sample of BaseEvaluator.evaluate (or the 100 k sequences
of BASELINE config 5) run as one batched decode.  Metric maths / MIDI round trips of the evaluators stay out of
scope (they need magenta / pretty_midi)."""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Normal

from . import ops
from ._lib import LIB, stream_ptr
from .ops import _p

EVENT_DIMS, RHYTHM_DIMS, NOTE_DIMS, CHROMA_DIMS = 342, 3, 16, 24


def convert_to_one_hot(input, dims):
    """test_class.py:34-41: (B,T) or (T,) int64 ids -> one-hot fp32 on the device."""
    return ops.ids_to_onehot(input.cuda(), dims)


def repar(mu, stddev, sigma=1):
    """test_class.py:53-56: eps drawn on the CPU default generator like the reference, z = mu + stddev * eps."""
    eps = Normal(0, sigma).sample(sample_shape=stddev.size()).to(stddev.device)
    return mu + stddev * eps


def clean_tokens(tokens: torch.Tensor):
    """Batched clean_output on the device: tokens (B, steps) int64 -> (start, length) int32 of the kept span of
    each row (leading / trailing zeros trimmed, cut at the first EOS = 1)."""
    tokens = tokens.long().contiguous()
    B, S = tokens.shape
    start = torch.empty(B, dtype=torch.int32, device=tokens.device)
    length = torch.empty(B, dtype=torch.int32, device=tokens.device)
    LIB.call("fn_clean_tokens", _p(tokens), B, S, _p(start), _p(length), stream_ptr(tokens.device))
    return start, length


def clean_output(out):
    """test_class.py:44-50 for one sequence: log-probs (1, steps, V) or (steps, V) -> numpy token array."""
    toks = out.argmax(-1).reshape(1, -1)
    start, length = clean_tokens(toks)
    s, l = int(start[0]), int(length[0])
    return toks[0, s:s + l].cpu().numpy()


def clean_outputs(tokens: torch.Tensor):
    """List of numpy token arrays, one per row of a batched decode."""
    start, length = clean_tokens(tokens)
    t, s, l = tokens.cpu().numpy(), start.cpu().numpy(), length.cpu().numpy()
    return [t[i, s[i]:s[i] + l[i]] for i in range(t.shape[0])]


@torch.no_grad()
def shift(model, d, r, n, c, target_z_value, attr="rhythm", steps=100, return_logp=True):
    """RhythmEvaluator.shift / NoteEvaluator.shift (test_class.py:233-254, 282-303) for a BATCH: forward, re-draw z
    from the returned distributions, overwrite latent dim 0 of the chosen attribute with `target_z_value`
    (scalar or (B,) tensor), switch the model to eval (as the reference does, permanently) and decode `steps`
    tokens greedily.  d, r, n: (B,T) int64 ids; c: (B,24).  Returns (out | tokens, original z[:,0])."""
    d_oh = convert_to_one_hot(d, EVENT_DIMS)
    r_oh = convert_to_one_hot(r, RHYTHM_DIMS)
    n_oh = convert_to_one_hot(n, NOTE_DIMS)
    res = model(d_oh, r_oh, n_oh, c.cuda().float())
    dis_r, dis_n = res[1]
    z_r, z_n = repar(dis_r.mean, dis_r.stddev), repar(dis_n.mean, dis_n.stddev)
    z_sel = z_r if attr == "rhythm" else z_n
    z0 = z_sel[:, 0].clone()
    z_sel[:, 0] = target_z_value
    model.eval()
    z = torch.cat([z_r, z_n, c.cuda().float()], dim=1)
    if return_logp:
        return model.global_decoder(z, steps=steps), z0
    return model.decode_greedy(z, steps, return_logp=False)[1], z0


@torch.no_grad()
def arousal_transfer(model, d, c, lam=1.0, steps=300, sample=True):
    """arousal_transfer.ipynb cells 11/15/17, batched: shift vectors mu_lookup(1) - mu_lookup(0), encode, (r)sample,
    z + lam * shift, greedy decode.  d: (B,T) ids or one-hot; returns tokens (B, steps) int64."""
    model.eval()
    idx = torch.tensor([0, 1], device=next(model.parameters()).device)
    mr, mn = model.mu_r_lookup(idx), model.mu_n_lookup(idx)
    dis_r, dis_n = model.encode(d)
    z_r = dis_r.rsample() if sample else dis_r.mean
    z_n = dis_n.rsample() if sample else dis_n.mean
    z = torch.cat([z_r + lam * (mr[1] - mr[0]), z_n + lam * (mn[1] - mn[0]), c.cuda().float()], dim=1)
    return model.decode_greedy(z, steps, return_logp=False)[1]
