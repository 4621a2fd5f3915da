"""Loss terms and the train / evaluate step shared by the two trainer mirrors
(reference trainer_gmm.py:109-293 and trainer.py:87-186).  All arithmetic on tensors goes
through the CUDA library (fadernets_b200.ops); only scalar bookkeeping is Python."""
from __future__ import annotations

import math

import numpy as np
import torch

from .ops import GmKlFn, LatentRegFn, NllMeanFn, StdKlFn
from .optim import FusedAdam


def beta_anneal(step, beta):
    """trainer_gmm.py:125-128 / trainer.py:93-96 (negative between steps 1000 and 9999, as the reference)."""
    return 0 if step < 1000 else min((step - 10000) / 10000 * beta, beta)


def reconstruction_terms(out, d, r_out, r, n_out, n):
    """CE_X, CE_R, CE_N = mean NLL over all B*T positions, no ignore_index (trainer_gmm.py:131-136)."""
    dev = out.device
    return (NllMeanFn.apply(out, d.to(dev)), NllMeanFn.apply(r_out, r.to(dev)), NllMeanFn.apply(n_out, n.to(dev)))


def gm_loss(model, out, d, r_out, r, n_out, n, dis, qy_x_out, logLogit_out, step, beta, is_supervised, y_label):
    """trainer_gmm.py:109-196 -> (loss, CE_X, CE_R, CE_N, kld_lat_r, kld_lat_n, kld_cls_r, kld_cls_n)."""
    beta0 = beta_anneal(step, beta)
    CE_X, CE_R, CE_N = reconstruction_terms(out, d, r_out, r, n_out, n)
    CE = 5 * CE_X + CE_R + CE_N
    dev = out.device
    mode = 1 if is_supervised else 0
    y = y_label.to(dev).long() if is_supervised else None
    terms = []
    for a, dist, qy, ll in (("r", dis[0], qy_x_out[0], logLogit_out[0]), ("n", dis[1], qy_x_out[1], logLogit_out[1])):
        mu_l, lv_l = getattr(model, f"mu_{a}_lookup").weight, getattr(model, f"logvar_{a}_lookup").weight
        terms.append(GmKlFn.apply(dist.mean, dist.stddev, mu_l, lv_l, qy, ll, y, mode))
    (lat_r, cls_r, clf_r), (lat_n, cls_n, clf_n) = terms[0].unbind(0), terms[1].unbind(0)
    if not is_supervised:
        loss = CE + beta0 * (lat_r + lat_n + cls_r + cls_n)
    else:
        loss = CE + beta0 * (lat_r + lat_n) + (clf_r + clf_n)
    return loss, CE_X, CE_R, CE_N, lat_r, lat_n, cls_r, cls_n


def vae_loss(out, d, r_out, r, n_out, n, dis, step, beta):
    """trainer.py:87-114 -> (loss, CE_X, CE_R, CE_N); KLD to N(0,1), mean over B*Z per latent."""
    beta0 = beta_anneal(step, beta)
    CE_X, CE_R, CE_N = reconstruction_terms(out, d, r_out, r, n_out, n)
    KLD = 0
    for dist in dis:
        KLD = KLD + StdKlFn.apply(dist.mean, dist.stddev)
    return 5 * CE_X + CE_R + CE_N + beta0 * KLD, CE_X, CE_R, CE_N


def _attr(a, dev):
    if torch.is_tensor(a):
        return a.to(dev, dtype=torch.float64)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)


def latent_reg(z_out, r, n):
    """trainer_gmm.py:199-217 / trainer.py:117-132 (Pati et al. 2019), latent dim 0 only."""
    z_r, z_n = z_out
    dev = z_r.device
    from . import parallel
    if parallel.global_latent_reg_enabled():           # pairs over the GLOBAL batch (one small all-gather per latent)
        return parallel.LatentRegGlobalFn.apply(z_r, _attr(r, dev)), parallel.LatentRegGlobalFn.apply(z_n, _attr(n, dev))
    return LatentRegFn.apply(z_r, _attr(r, dev)), LatentRegFn.apply(z_n, _attr(n, dev))


def optimise(model, optimizer, loss):
    """loss.backward(); clip_grad_norm_(params, 1); optimizer.step()  (trainer_gmm.py:249-251)."""
    scale = getattr(getattr(optimizer, "grad_sync", None), "loss_scale", 1.0)
    (loss if scale == 1.0 else loss * scale).backward()        # data parallel: the 1/world of the gradient mean, folded into backward
    if isinstance(optimizer, FusedAdam):
        optimizer.step()                       # global-norm clip (max_norm=1) is fused into the update
    else:
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimizer.step()


def to_floats(*scalars):
    """One device->host transfer instead of the reference's 8 `.item()` syncs (+ the index-range verdict of this step)."""
    from . import ops
    vals = torch.stack([s.reshape(()) for s in scalars])
    bad = ops._bad_counter(vals.device)
    out = torch.cat([vals.float(), bad.float()]).tolist()
    if out[-1]:
        bad.zero_()
        raise IndexError(f"fadernets_b200: {int(out[-1])} token / target / label index(es) out of range "
                         "(the reference's nll_loss / Embedding raise here too)")
    return tuple(out[:-1])
