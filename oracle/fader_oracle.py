"""CPU oracle for the GM-VAE / vanilla-VAE hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's
algorithm for the train step and the decode path.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``music-fader-nets_b200/``
imports it, and the product path fails loudly when its CUDA library is missing.

Parity status: the reference has no tests / golden vectors of its own
(SURVEY.md section 4), so this oracle is pinned against the *reference code itself*,
imported in the build container from /root/reference by ``oracle/gen_golden.py``
(fixtures committed under ``tests/golden/``) and live in
``tests/test_oracle_vs_reference.py`` (incl. the shipped ``params/*.pt`` checkpoints).

Every function cites the reference lines it restates (paths relative to
/root/reference).  The arithmetic itself lives in PyTorch (un-pinned dependency of
the reference): GRU gate equations torch ``nn/modules/rnn.py`` (GRU docstring),
Normal-Normal KL ``distributions/kl.py`` (_kl_normal_normal), gradient clipping
``nn/utils/clip_grad.py``, Adam ``optim/adam.py`` -- restated explicitly below.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EVENT_DIMS, RHYTHM_DIMS, NOTE_DIMS, CHROMA_DIMS = 342, 3, 16, 24  # trainer_gmm.py:35-38
START_TOKEN = 341                                                   # gmm_model.py:120-121

Tensor = torch.Tensor
Weights = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# parameter set                                                                #
# --------------------------------------------------------------------------- #
def param_shapes(hidden: int, z: int, variant: str = "gmvae", n_component: int = 2,
                 roll: int = EVENT_DIMS, rhythm: int = RHYTHM_DIMS, note: int = NOTE_DIMS,
                 chroma: int = CHROMA_DIMS) -> Dict[str, Tuple[int, ...]]:
    """state_dict names -> shapes, in the reference's registration order.

    gmm_model.py:32-71 (GM-VAE) / model_v2.py:25-60 (vanilla).  The two classes
    register ``c_r/c_n`` at different positions; order matters only for
    ``state_dict()`` iteration order, which we keep.
    """
    H, Z, V = hidden, z, roll
    G = 2 * Z + 24                                   # gmm_model.py:52-55 (cdtl_dims = 24)
    s: Dict[str, Tuple[int, ...]] = {}

    def gru(name, inp, bi):
        for sfx in ([""] + (["_reverse"] if bi else [])):
            s[f"{name}.weight_ih_l0{sfx}"] = (3 * H, inp)
            s[f"{name}.weight_hh_l0{sfx}"] = (3 * H, H)
            s[f"{name}.bias_ih_l0{sfx}"] = (3 * H,)
            s[f"{name}.bias_hh_l0{sfx}"] = (3 * H,)

    def lin(name, inp, out):
        s[f"{name}.weight"] = (out, inp)
        s[f"{name}.bias"] = (out,)

    def cell(name, inp):
        s[f"{name}.weight_ih"] = (3 * H, inp)
        s[f"{name}.weight_hh"] = (3 * H, H)
        s[f"{name}.bias_ih"] = (3 * H,)
        s[f"{name}.bias_hh"] = (3 * H,)

    for g in ("gru_r", "gru_n", "gru_c"):
        gru(g, V, True)
    if variant == "gmvae":
        lin("c_r", Z, 3); lin("c_n", Z, 3)
    gru("gru_d_r", Z + rhythm, False); gru("gru_d_n", Z + note, False); gru("gru_d_c", Z + chroma, False)
    if variant == "vae":
        lin("c_r", Z, 3); lin("c_n", Z, 3)
    for a in ("r", "n", "c"):
        lin(f"mu_{a}", 2 * H, Z); lin(f"var_{a}", 2 * H, Z)
    lin("linear_init_global", G, H)
    cell("grucell_g", G + V)
    cell("grucell_g_2", H)
    for a in ("r", "n", "c"):
        lin(f"linear_init_{a}", Z, H)
    lin("linear_out_r", H, rhythm); lin("linear_out_n", H, note)
    lin("linear_out_c", Z, chroma); lin("linear_out_g", H, V)
    if variant == "gmvae":
        s["mu_r_lookup.weight"] = (n_component, Z)
        s["mu_n_lookup.weight"] = (n_component, Z)
        s["logvar_r_lookup.weight"] = (n_component, Z)
        s["logvar_n_lookup.weight"] = (n_component, Z)
    return s


# parameters that never receive a gradient in the reference (probed):
DEAD_PREFIXES = ("gru_c.", "gru_d_c.", "c_r.", "c_n.", "mu_c.", "var_c.", "linear_init_c.", "linear_out_c.")
FROZEN = ("logvar_r_lookup.weight", "logvar_n_lookup.weight")       # gmm_model.py:175,182


def is_live(name: str) -> bool:
    return not name.startswith(DEAD_PREFIXES) and name not in FROZEN


def init_weights(hidden: int, z: int, variant: str = "gmvae", n_component: int = 2,
                 seed: int = 0, dtype=torch.float32) -> Weights:
    """Deterministic synthetic weights (NOT the reference's init order -- tests that
    need reference-initialised weights take the reference's own state_dict).
    uniform(+-1/sqrt(H)) like torch's RNN/Linear defaults; logvar lookups = -4
    (gmm_model.py:172-174: log(exp(-2)**2))."""
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / math.sqrt(hidden)
    w: Weights = {}
    for name, shp in param_shapes(hidden, z, variant, n_component).items():
        if name.startswith("logvar_"):
            w[name] = torch.full(shp, math.log(math.exp(-2) ** 2), dtype=dtype)
        else:
            w[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype)
    return w


# --------------------------------------------------------------------------- #
# GRU building blocks (torch nn/modules/rnn.py GRU docstring; gate order r,z,n) #
# --------------------------------------------------------------------------- #
def gru_cell(gi: Tensor, h: Tensor, w_hh: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step given the input-side pre-activations ``gi = W_ih x + b_ih``.

    r = sigma(gi_r + gh_r); z = sigma(gi_z + gh_z); n = tanh(gi_n + r * gh_n);
    h' = (1 - z) * n + z * h,   gh = W_hh h + b_hh.
    """
    H = h.shape[-1]
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[..., :H] + gh[..., :H])
    zt = torch.sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H])
    n = torch.tanh(gi[..., 2 * H:] + r * gh[..., 2 * H:])
    return (1 - zt) * n + zt * h


def gru_seq(gi: Tensor, h0: Tensor, w_hh: Tensor, b_hh: Tensor, reverse: bool = False) -> Tensor:
    """Run a GRU over time.  gi: (B,T,3H) precomputed input side.  Returns all hidden
    states (B,T,H) aligned with input time (nn.GRU ``output`` semantics)."""
    B, T, _ = gi.shape
    h = h0
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h = gru_cell(gi[:, t], h, w_hh, b_hh)
        outs[t] = h
    return torch.stack(outs, 1)


def input_side(w_ih: Tensor, b_ih: Tensor, ids: Optional[Tensor], dense: Optional[Tensor],
               dense_onehot: bool) -> Tensor:
    """gi for a one-hot (token id) input.  ``one_hot @ W_ih^T == W_ih^T[ids]`` exactly in
    floating point (a single 1.0*w product per output, plus exact zeros), so both
    forms give identical numbers; ``dense_onehot`` reproduces the reference's cost
    (a dense GEMM on the one-hot tensor, gmm_model.py:84) for CPU-baseline timing."""
    if dense_onehot:
        x = F.one_hot(ids, w_ih.shape[1]).to(w_ih.dtype) if dense is None else dense
        return x @ w_ih.t() + b_ih
    return w_ih.t()[ids] + b_ih


# --------------------------------------------------------------------------- #
# model pieces                                                                 #
# --------------------------------------------------------------------------- #
def encode(w: Weights, d_ids: Tensor, dense_onehot: bool = False):
    """gmm_model.py:82-98 / model_v2.py:81-97.

    Two bidirectional GRUs over the event one-hots; final states concatenated
    [h_fwd(T-1) | h_bwd(0)] (``transpose_(0,1).view(B,-1)`` of h_n (2,B,H));
    mu = Linear, scale = exp(Linear) (used as the Normal *scale*, :86,:93).
    Returns (mu_r, scale_r, mu_n, scale_n)."""
    B, T = d_ids.shape
    out = []
    for a in ("r", "n"):
        hs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            w_ih, w_hh = w[f"gru_{a}.weight_ih_l0{sfx}"], w[f"gru_{a}.weight_hh_l0{sfx}"]
            b_ih, b_hh = w[f"gru_{a}.bias_ih_l0{sfx}"], w[f"gru_{a}.bias_hh_l0{sfx}"]
            gi = input_side(w_ih, b_ih, d_ids, None, dense_onehot)
            h0 = torch.zeros(B, w_hh.shape[1], dtype=w_hh.dtype)
            o = gru_seq(gi, h0, w_hh, b_hh, reverse=rev)
            hs.append(o[:, 0] if rev else o[:, T - 1])
        hcat = torch.cat(hs, 1)
        mu = hcat @ w[f"mu_{a}.weight"].t() + w[f"mu_{a}.bias"]
        scale = torch.exp(hcat @ w[f"var_{a}.weight"].t() + w[f"var_{a}.bias"])
        out += [mu, scale]
    return tuple(out)


def approx_qy_x(z: Tensor, mu_lookup: Tensor, logvar_lookup: Tensor):
    """gmm_model.py:194-218.  logLogit[b,k] = -0.5*sum_d((z-mu_k)^2/exp(lv_k) + lv_k + ln 2pi)
    + ln(1/K);  qy_x = softmax_k(logLogit).  (exp(logvar) is the *variance* here.)"""
    K = mu_lookup.shape[0]
    d = z[:, None, :] - mu_lookup[None]
    llh = -0.5 * (d * d / torch.exp(logvar_lookup)[None] + logvar_lookup[None] + math.log(2 * math.pi))
    logLogit = llh.sum(-1) + math.log(1.0 / K)
    return logLogit, torch.softmax(logLogit, 1)


def sub_decoders(w: Weights, r_ids: Tensor, z_r: Tensor, n_ids: Tensor, z_n: Tensor,
                 dense_onehot: bool = False):
    """gmm_model.py:100-117 / model_v2.py:99-116.

    input_t = [onehot(attr_t), z] ; h0 = linear_init(z); uni-GRU; Linear(H->dims);
    **log_softmax over dim=1, the TIME axis** (:110,:115)."""
    outs = []
    for a, ids, z, dims in (("r", r_ids, z_r, RHYTHM_DIMS), ("n", n_ids, z_n, NOTE_DIMS)):
        w_ih, w_hh = w[f"gru_d_{a}.weight_ih_l0"], w[f"gru_d_{a}.weight_hh_l0"]
        b_ih, b_hh = w[f"gru_d_{a}.bias_ih_l0"], w[f"gru_d_{a}.bias_hh_l0"]
        if dense_onehot:
            T = ids.shape[1]
            x = torch.cat([F.one_hot(ids, dims).to(z.dtype), z[:, None, :].expand(-1, T, -1)], -1)
            gi = x @ w_ih.t() + b_ih
        else:
            gi = w_ih[:, :dims].t()[ids] + (z @ w_ih[:, dims:].t() + b_ih)[:, None, :]
        h0 = z @ w[f"linear_init_{a}.weight"].t() + w[f"linear_init_{a}.bias"]
        hs = gru_seq(gi, h0, w_hh, b_hh)
        logits = hs @ w[f"linear_out_{a}.weight"].t() + w[f"linear_out_{a}.bias"]
        outs.append(torch.log_softmax(logits, 1))
    return tuple(outs)


def global_decoder(w: Weights, z: Tensor, steps: int, teacher_ids: Optional[Tensor] = None,
                   dense_onehot: bool = False):
    """gmm_model.py:119-149 / model_v2.py:118-143.

    out_0 = onehot(341); hx0 = linear_init_global(z); per step i:
      hx0 = GRUCell_g([out, z], hx0); (i == 0: hx1 <- the *new* hx0);
      hx1 = GRUCell_g2(hx0, hx1); out = log_softmax(linear_out_g(hx1)) over the vocab;
      next input: training -> teacher one-hot x[:, i] (always: p < eps = 100, :139-142),
                  eval     -> one-hot of the first-max index (_sampling, :73-80).
    ``teacher_ids`` given => training-mode teacher forcing; None => greedy (eval).
    Returns (log-probs (B,steps,V), greedy/teacher token ids fed back (B,steps))."""
    V = EVENT_DIMS
    B = z.shape[0]
    w_ih1, w_hh1 = w["grucell_g.weight_ih"], w["grucell_g.weight_hh"]
    b_ih1, b_hh1 = w["grucell_g.bias_ih"], w["grucell_g.bias_hh"]
    w_ih2, w_hh2 = w["grucell_g_2.weight_ih"], w["grucell_g_2.weight_hh"]
    b_ih2, b_hh2 = w["grucell_g_2.bias_ih"], w["grucell_g_2.bias_hh"]
    w_o, b_o = w["linear_out_g.weight"], w["linear_out_g.bias"]
    hx0 = z @ w["linear_init_global.weight"].t() + w["linear_init_global.bias"]
    hx1 = None
    tok = torch.full((B,), START_TOKEN, dtype=torch.long)
    zproj = z @ w_ih1[:, V:].t() + b_ih1                       # time-invariant part of W_ih [out, z]
    w_tok = w_ih1[:, :V].t()
    outs, fed = [], []
    for i in range(steps):
        if dense_onehot:
            gi1 = torch.cat([F.one_hot(tok, V).to(z.dtype), z], 1) @ w_ih1.t() + b_ih1
        else:
            gi1 = w_tok[tok] + zproj
        hx0 = gru_cell(gi1, hx0, w_hh1, b_hh1)
        if i == 0:
            hx1 = hx0
        hx1 = gru_cell(hx0 @ w_ih2.t() + b_ih2, hx1, w_hh2, b_hh2)
        lp = torch.log_softmax(hx1 @ w_o.t() + b_o, 1)
        outs.append(lp)
        tok = teacher_ids[:, i] if teacher_ids is not None else lp.max(1)[1]
        fed.append(tok)
    return torch.stack(outs, 1), torch.stack(fed, 1)


def forward(w: Weights, variant: str, d_ids: Tensor, r_ids: Tensor, n_ids: Tensor, c: Tensor,
            eps_r: Tensor, eps_n: Tensor, training: bool = True, dense_onehot: bool = False):
    """gmm_model.py:220-259 / model_v2.py:145-171 with the noise passed in explicitly
    (the reference draws eps_r then eps_n from the CPU default generator, :229-235)."""
    mu_r, s_r, mu_n, s_n = encode(w, d_ids, dense_onehot)
    z_r = mu_r + s_r * eps_r
    z_n = mu_n + s_n * eps_n
    res = dict(mu_r=mu_r, scale_r=s_r, mu_n=mu_n, scale_n=s_n, z_r=z_r, z_n=z_n)
    if variant == "gmvae":
        for a, z in (("r", z_r), ("n", z_n)):
            ll, q = approx_qy_x(z, w[f"mu_{a}_lookup.weight"], w[f"logvar_{a}_lookup.weight"])
            res[f"logLogit_{a}"], res[f"qy_x_{a}"], res[f"y_{a}"] = ll, q, q.max(1)[1]
    res["r_out"], res["n_out"] = sub_decoders(w, r_ids, z_r, n_ids, z_n, dense_onehot)
    zc = torch.cat([z_r, z_n, c], 1)
    res["out"], res["fed"] = global_decoder(w, zc, d_ids.shape[1], d_ids if training else None, dense_onehot)
    return res


# --------------------------------------------------------------------------- #
# losses                                                                       #
# --------------------------------------------------------------------------- #
def beta_anneal(step: int, beta: float) -> float:
    """trainer_gmm.py:125-128 -- note: negative for 1000 <= step < 10000."""
    return 0.0 if step < 1000 else min((step - 10000) / 10000 * beta, beta)


def kl_normal(mu_q, s_q, mu_p, s_p):
    """torch distributions/kl.py _kl_normal_normal: 0.5*(rho + t1 - 1 - ln rho),
    rho=(s_q/s_p)^2, t1=((mu_q-mu_p)/s_p)^2  (KL(q || p))."""
    rho = (s_q / s_p) ** 2
    t1 = ((mu_q - mu_p) / s_p) ** 2
    return 0.5 * (rho + t1 - 1 - torch.log(rho))


def nll_mean(logp: Tensor, tgt: Tensor) -> Tensor:
    """F.nll_loss(..., reduction='mean') without ignore_index (trainer_gmm.py:131-136)."""
    return -logp.reshape(-1, logp.shape[-1]).gather(1, tgt.reshape(-1, 1)).mean()


def loss_gmvae(w: Weights, res, d, r, n, step: int, beta: float, y_label: Optional[Tensor] = None):
    """trainer_gmm.py:109-196.  Returns the 8-tuple
    (loss, CE_X, CE_R, CE_N, kld_lat_r, kld_lat_n, kld_cls_r, kld_cls_n)."""
    beta0 = beta_anneal(step, beta)
    ce_x, ce_r, ce_n = nll_mean(res["out"], d), nll_mean(res["r_out"], r), nll_mean(res["n_out"], n)
    ce = 5 * ce_x + ce_r + ce_n
    K = w["mu_r_lookup.weight"].shape[0]
    out = {}
    if y_label is None:
        for a in ("r", "n"):
            tot = 0
            for k in range(K):
                # scale of p(z|y=k) is exp(logvar) (NOT exp(0.5*logvar)) -- trainer_gmm.py:156-157
                kl = kl_normal(res[f"mu_{a}"], res[f"scale_{a}"], w[f"mu_{a}_lookup.weight"][k],
                               torch.exp(w[f"logvar_{a}_lookup.weight"][k])).mean(-1)
                tot = tot + (kl * res[f"qy_x_{a}"][:, k]).mean()
            out[f"lat_{a}"] = tot
            ent = (res[f"qy_x_{a}"] * torch.log_softmax(res[f"logLogit_{a}"], 1)).mean(1)   # :170-171
            out[f"cls_{a}"] = (ent - math.log(1.0 / K)).mean()
        loss = ce + beta0 * (out["lat_r"] + out["lat_n"] + out["cls_r"] + out["cls_n"])
    else:
        clf = 0
        for a in ("r", "n"):
            kl = kl_normal(res[f"mu_{a}"], res[f"scale_{a}"], w[f"mu_{a}_lookup.weight"][y_label],
                           torch.exp(w[f"logvar_{a}_lookup.weight"][y_label])).mean(-1)
            out[f"lat_{a}"] = kl.mean()
            out[f"cls_{a}"] = torch.zeros((), dtype=kl.dtype)
            # nn.CrossEntropyLoss applied to the *probabilities* qy_x (trainer_gmm.py:192-193)
            clf = clf + nll_mean(torch.log_softmax(res[f"qy_x_{a}"], 1), y_label)
        loss = ce + beta0 * (out["lat_r"] + out["lat_n"]) + clf
    return loss, ce_x, ce_r, ce_n, out["lat_r"], out["lat_n"], out["cls_r"], out["cls_n"]


def loss_vae(res, d, r, n, step: int, beta: float):
    """trainer.py:87-114: KLD = sum over the two latents of mean_{B,Z} KL(q || N(0,1))."""
    beta0 = beta_anneal(step, beta)
    ce_x, ce_r, ce_n = nll_mean(res["out"], d), nll_mean(res["r_out"], r), nll_mean(res["n_out"], n)
    kld = 0
    for a in ("r", "n"):
        mu, s = res[f"mu_{a}"], res[f"scale_{a}"]
        kld = kld + kl_normal(mu, s, torch.zeros_like(mu), torch.ones_like(s)).mean()
    return 5 * ce_x + ce_r + ce_n + beta0 * kld, ce_x, ce_r, ce_n


def latent_reg(z_r: Tensor, z_n: Tensor, r_density, n_density):
    """trainer_gmm.py:199-217 / trainer.py:117-132 (Pati et al. 2019):
    D_attr = outer difference of the densities computed in float64 on the host then cast
    to float; l = mean_{BxB}(tanh(z0_i - z0_j) - sign(D_attr_ij))^2 on latent dim 0."""
    outs = []
    for z, dens in ((z_r, r_density), (z_n, n_density)):
        dens = np.asarray(dens, dtype=np.float64)
        sgn = torch.sign(torch.from_numpy(np.subtract.outer(dens, dens)).float()).to(z.dtype)
        dz = z[:, 0].reshape(-1, 1) - z[:, 0]
        outs.append(((torch.tanh(dz) - sgn) ** 2).mean())
    return tuple(outs)


# --------------------------------------------------------------------------- #
# train step (autograd for the derivative; clip + Adam restated)               #
# --------------------------------------------------------------------------- #
class AdamState:
    """optim.Adam(lr, betas=(0.9,0.999), eps=1e-8) state (trainer_gmm.py:52)."""

    def __init__(self, w: Weights):
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in w.items() if is_live(k)}
        self.v = {k: torch.zeros_like(v) for k, v in w.items() if is_live(k)}


def clip_and_adam(w: Weights, grads: Dict[str, Tensor], st: AdamState, lr: float,
                  max_norm: float = 1.0) -> float:
    """clip_grad_norm_(params, 1) (trainer_gmm.py:250; torch clip_grad.py: total L2 norm
    over all grads, coef = max_norm/(norm+1e-6) clamped to <= 1) then Adam.step()
    (bias-corrected: step = lr/(1-b1^t); denom = sqrt(v)/sqrt(1-b2^t) + eps)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    coef = min(float(max_norm / (total + 1e-6)), 1.0)
    st.t += 1
    b1, b2, eps = 0.9, 0.999, 1e-8
    bc1, bc2 = 1 - b1 ** st.t, 1 - b2 ** st.t
    for k, g in grads.items():
        g = g * coef
        st.m[k].mul_(b1).add_(g, alpha=1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = st.v[k].sqrt() / math.sqrt(bc2) + eps
        w[k] = w[k] - (lr / bc1) * st.m[k] / denom
    return float(total)


def loss_and_grads(w: Weights, variant: str, batch, eps_r, eps_n, step: int, beta: float,
                   y_label=None, dense_onehot: bool = False):
    """Forward + all loss terms + gradients w.r.t. every live parameter
    (trainer_gmm.py:223-249 / trainer.py:138-156 up to ``loss.backward()``)."""
    d, r, n, c, r_density, n_density = batch
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in w.items() if is_live(k)}
    ww = {**w, **leaves}
    res = forward(ww, variant, d, r, n, c, eps_r, eps_n, True, dense_onehot)
    if variant == "gmvae":
        terms = loss_gmvae(ww, res, d, r, n, step, beta, y_label)
    else:
        terms = loss_vae(res, d, r, n, step, beta)
    l_r, l_n = latent_reg(res["z_r"], res["z_n"], r_density, n_density)
    loss = terms[0] + l_r + l_n
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    scalars = dict(loss=loss.detach(), CE_X=terms[1].detach(), CE_R=terms[2].detach(), CE_N=terms[3].detach(),
                   l_r=l_r.detach(), l_n=l_n.detach())
    if variant == "gmvae":
        scalars.update(kld_lat_r=terms[4].detach(), kld_lat_n=terms[5].detach(),
                       kld_cls_r=terms[6].detach(), kld_cls_n=terms[7].detach())
    return scalars, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in res.items()}


def train_step(w: Weights, st: AdamState, variant: str, batch, eps_r, eps_n, step: int,
               beta: float, lr: float, y_label=None, dense_onehot: bool = False):
    """One full ``train()`` call (trainer_gmm.py:220-258 / trainer.py:135-162).
    Mutates ``w`` (dict entries replaced) and ``st``.  Returns (scalars, grad_norm)."""
    scalars, grads, _ = loss_and_grads(w, variant, batch, eps_r, eps_n, step, beta, y_label, dense_onehot)
    norm = clip_and_adam(w, grads, st, lr)
    return scalars, norm


# --------------------------------------------------------------------------- #
# synthetic data                                              #
# --------------------------------------------------------------------------- #
def synth_batch(B: int, T: int, seed: int = 0, pad_tail: bool = False):
    """Seeded synthetic event-token batch in the dataset's tuple layout
    (ptb_v2.py:400-436): d in [2,342), r in {0,1,2}, n in [0,16), c ~ U[0,1) (B,24),
    r_density = mean(r == 1), n_density = mean(n) as float64 (ptb_v2.py:421-422)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randint(2, EVENT_DIMS, (B, T), generator=g)
    if pad_tail:
        k = max(1, T // 8)
        d[:, T - k] = 1
        d[:, T - k + 1:] = 0
    r = torch.randint(0, RHYTHM_DIMS, (B, T), generator=g)
    n = torch.randint(0, NOTE_DIMS, (B, T), generator=g)
    c = torch.rand(B, CHROMA_DIMS, generator=g)
    r_density = (r == 1).double().mean(1).numpy()
    n_density = n.double().mean(1).numpy()
    return d, r, n, c, r_density, n_density


def draw_eps(B: int, Z: int, T: int, training: bool = True):
    """Consume the CPU default generator exactly as one reference forward does
    (gmm_model.py:229-235 then T x torch.rand(1) in the decoder loop, :140)."""
    eps_r = torch.normal(torch.zeros(B, Z), torch.ones(B, Z))
    eps_n = torch.normal(torch.zeros(B, Z), torch.ones(B, Z))
    if training:
        torch.rand(T)
    return eps_r, eps_n


# --------------------------------------------------------------------------- #
# sibling models on the same blocks: model_v2.py:174-586         #
# --------------------------------------------------------------------------- #
SIBLING_DEAD = ("c_r.", "c_n.")                      # registered by the reference, never used in forward (model_v2.py:305-306, 458-459)


def sibling_is_live(name: str) -> bool:
    return not name.startswith(SIBLING_DEAD)


def _bi_encoder(w: Weights, gru: str, d_ids: Tensor, extra: Optional[Tensor] = None):
    """One bidirectional GRU over the event one-hots (plus `extra` (B,E): columns appended to every step's input,
    model_v2.py:338-341) -> [h_fwd(T-1) | h_bwd(0)] -> mu = Linear, scale = exp(Linear) (model_v2.py:229-233, 343-347, 497-502)."""
    B, T = d_ids.shape
    V = EVENT_DIMS
    hs = []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih, w_hh = w[f"{gru}.weight_ih_l0{sfx}"], w[f"{gru}.weight_hh_l0{sfx}"]
        b_ih, b_hh = w[f"{gru}.bias_ih_l0{sfx}"], w[f"{gru}.bias_hh_l0{sfx}"]
        gi = w_ih[:, :V].t()[d_ids] + b_ih
        if extra is not None:
            gi = gi + (extra @ w_ih[:, V:].t())[:, None, :]
        o = gru_seq(gi, torch.zeros(B, w_hh.shape[1], dtype=w_hh.dtype), w_hh, b_hh, reverse=rev)
        hs.append(o[:, 0] if rev else o[:, T - 1])
    hcat = torch.cat(hs, 1)
    return hcat @ w["mu.weight"].t() + w["mu.bias"], torch.exp(hcat @ w["var.weight"].t() + w["var.bias"])


def forward_sibling(w: Weights, kind: str, d_ids: Tensor, c: Tensor, r_density: Tensor, n_density: Tensor, eps: Tensor,
                    mask_r: Optional[Tensor] = None, mask_n: Optional[Tensor] = None, training: bool = True):
    """forward() of MusicAttrSingleVAE (model_v2.py:263-285), MusicAttrCVAE (:400-423) and MusicAttrFaderNets (:560-586)
    with the noise (and, FaderNets, the two dropout masks incl. their 1/(1-p) scale) passed in explicitly.
    r_density / n_density: (B,1) as the trainers pass them (trainer_cvae.py:122-125)."""
    T = d_ids.shape[1]
    res = {}
    if kind == "singlevae":
        mu, s = _bi_encoder(w, "gru", d_ids)
        z = mu + s * eps
        zc = torch.cat([z, c], 1)
    elif kind == "cvae":
        mu, s = _bi_encoder(w, "gru_e", d_ids, torch.cat([r_density, n_density], 1))
        z = mu + s * eps
        zc = torch.cat([z, r_density, n_density], 1)
    else:
        mu, s = _bi_encoder(w, "gru_e", d_ids)
        z = mu + s * eps
        # gradient reversal is the identity in forward (ReverseLayerF, :426-435); dropout(relu(linear)) (:574-575)
        rz = GradReverse.apply(z)
        r_out = torch.relu(rz @ w["discriminator_r.weight"].t() + w["discriminator_r.bias"])
        n_out = torch.relu(rz @ w["discriminator_n.weight"].t() + w["discriminator_n.bias"])
        res["r_out"] = r_out * mask_r if mask_r is not None else r_out
        res["n_out"] = n_out * mask_n if mask_n is not None else n_out
        zc = torch.cat([z, r_density, n_density], 1)
    res.update(mu=mu, scale=s, z=zc, z_lat=z)
    res["out"], res["fed"] = global_decoder(w, zc, T, d_ids if training else None)
    return res


class GradReverse(torch.autograd.Function):
    """ReverseLayerF (model_v2.py:426-435): identity forward, negated gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg()


def loss_sibling(kind: str, res, d, r_density, n_density, step: int, beta: float):
    """trainer_singlevae.py:86-123, trainer_cvae.py:84-102, trainer_fader.py:84-110.  Returns (loss, named terms).
    Quirks kept: the single-VAE trainer weights the KL with `beta` (not the annealed beta0) and the CE with 5."""
    beta0 = beta_anneal(step, beta)
    ce_x = nll_mean(res["out"], d)
    kld = kl_normal(res["mu"], res["scale"], torch.zeros_like(res["mu"]), torch.ones_like(res["scale"])).mean()
    if kind == "singlevae":
        outs = []
        for col, dens in ((0, r_density), (1, n_density)):
            dens = np.asarray(dens, dtype=np.float64).reshape(-1)
            sgn = torch.sign(torch.from_numpy(np.subtract.outer(dens, dens)).float()).to(res["z"].dtype)
            dz = res["z"][:, col].reshape(-1, 1) - res["z"][:, col]
            outs.append(((torch.tanh(dz) - sgn) ** 2).mean())
        loss = 5 * ce_x + beta * kld + outs[0] + outs[1]
        return loss, dict(CE_X=ce_x, l_r=outs[0], l_n=outs[1])
    if kind == "cvae":
        return ce_x + beta0 * kld, dict(CE_X=ce_x)
    lmbda = min(step / 2000 * 1e-4, 1e-4)
    l_r = lmbda * ((res["r_out"].squeeze() - r_density.squeeze()) ** 2).mean()
    l_n = lmbda * ((res["n_out"].squeeze() - n_density.squeeze()) ** 2).mean()
    return ce_x + beta0 * kld + l_r + l_n, dict(CE_X=ce_x, l_adv_r=l_r, l_adv_n=l_n)


def sibling_loss_and_grads(w: Weights, kind: str, d, c, r_density, n_density, eps, step, beta, mask_r=None, mask_n=None):
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in w.items() if sibling_is_live(k)}
    ww = {**w, **leaves}
    rd = torch.as_tensor(np.asarray(r_density), dtype=torch.float32).reshape(-1, 1)
    nd = torch.as_tensor(np.asarray(n_density), dtype=torch.float32).reshape(-1, 1)
    res = forward_sibling(ww, kind, d, c, rd, nd, eps, mask_r, mask_n)
    loss, terms = loss_sibling(kind, res, d, rd if kind != "singlevae" else r_density, nd if kind != "singlevae" else n_density,
                               step, beta)
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    scal = dict(loss=loss.detach(), **{k: v.detach() for k, v in terms.items()})
    return scal, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in res.items()}


# --------------------------------------------------------------------------- #
# GLSR regulariser: trainer_glsr.py:118-229                     #
# --------------------------------------------------------------------------- #
GLSR_STEPS = 100                  # the regulariser decodes 100 teacher-forced steps (trainer_glsr.py:187,189,212,214)
GLSR_EPS = 1e-2
PLAYED_NOTES = (2, 90)            # tokens 2..89: MIDI note-on events (trainer_glsr.py:124-126)
TIME_SHIFTS = (180, 278)          # tokens 180..277: time shifts used as step separators (:132-134)


def glsr_note_density(logp: Tensor) -> Tensor:
    """approx_note_density (trainer_glsr.py:139-141): per sequence, the probability mass on the note-on tokens summed over time.
    The reference applies softmax to the decoder's log-probabilities (a re-normalisation of exp(logp))."""
    p = torch.softmax(logp, -1)
    return p[..., PLAYED_NOTES[0]:PLAYED_NOTES[1]].sum(-1).sum(1)                     # (B,)


def glsr_rhythm_density(logp: Tensor) -> Tensor:
    """approx_rhythm_density (trainer_glsr.py:143-171), vectorised over the batch with the reference's semantics AS WRITTEN:
    the running note mass `cur` always reads sequence 0 (`played_notes[0][i]`, :154); at a step whose time-shift mass is
    >= 0.9 a non-zero `cur` is flushed into `total` as 1 (`cur / cur`: value 1, zero gradient) if it exceeds 1e-2, else as
    itself; what is left after the last separator is dropped; density = total / (time-shift mass summed over time), and a
    density of exactly 0 is replaced by a constant."""
    p = torch.softmax(logp, -1)
    notes0 = p[0, :, PLAYED_NOTES[0]:PLAYED_NOTES[1]].sum(-1)                         # (S,)  sequence 0, whatever the row
    seps = p[..., TIME_SHIFTS[0]:TIME_SHIFTS[1]].sum(-1)                              # (B,S)
    B, S = seps.shape
    cur = torch.zeros(B, dtype=logp.dtype)
    total = torch.zeros(B, dtype=logp.dtype)
    for i in range(S):
        is_sep = seps[:, i].detach() >= 0.9
        cur = torch.where(is_sep, cur, cur + notes0[i])
        flush = is_sep & (cur.detach() != 0)
        piece = torch.where(cur.detach() > 1e-2, torch.ones_like(cur), cur)
        total = total + torch.where(flush, piece, torch.zeros_like(cur))
        cur = torch.where(flush, torch.zeros_like(cur), cur)
    dens = total / seps.sum(1)
    return torch.where(total.detach() != 0, dens, torch.zeros_like(dens))


def glsr_regulariser(w: Weights, z_r: Tensor, z_n: Tensor, c: Tensor, teacher_ids: Tensor, deltas_r: Tensor, deltas_n: Tensor):
    """latent_regularized_loss_function of trainer_glsr.py:118-229 (Hadjeres et al.'s GLSR as the reference implements it):
    finite differences of the two approximated attributes along latent dimension 0, through FOUR extra teacher-forced decodes
    of 100 steps, pushed towards a standard normal: l = mean(-log N(g; 0, 1)), g = (a(z+) - a(z-)) / (2 delta)."""
    def fd(attr, zs_plus, zs_minus, deltas):
        a_p = attr(global_decoder(w, torch.cat(zs_plus + [c], 1), GLSR_STEPS, teacher_ids[:, :GLSR_STEPS])[0])
        a_m = attr(global_decoder(w, torch.cat(zs_minus + [c], 1), GLSR_STEPS, teacher_ids[:, :GLSR_STEPS])[0])
        g = (a_p - a_m) / (2 * deltas)
        return (0.5 * g * g + 0.5 * math.log(2 * math.pi)).mean()

    def shifted(z, d):
        zz = z.clone()
        zz[:, 0] = zz[:, 0] + d
        return zz
    l_r = fd(glsr_rhythm_density, [shifted(z_r, deltas_r), z_n], [shifted(z_r, -deltas_r), z_n], deltas_r)
    l_n = fd(glsr_note_density, [z_r, shifted(z_n, deltas_n)], [z_r, shifted(z_n, -deltas_n)], deltas_n)
    return l_r, l_n


def draw_glsr_deltas(B: int):
    """The CPU-generator draws of one regulariser call, in the reference's order: deltas for z_r, the coin flips of its two
    decodes, deltas for z_n, the coin flips of its two decodes (trainer_glsr.py:179,187,189,204,212,214; model_v2.py:136)."""
    d_r = (1 + torch.rand(B)) * GLSR_EPS
    torch.rand(GLSR_STEPS); torch.rand(GLSR_STEPS)
    d_n = (1 + torch.rand(B)) * GLSR_EPS
    torch.rand(GLSR_STEPS); torch.rand(GLSR_STEPS)
    return d_r, d_n


def glsr_loss_and_grads(w: Weights, batch, eps_r, eps_n, deltas_r, deltas_n, step: int, beta: float):
    """Forward + loss_function + GLSR regulariser + gradients of reference trainer_glsr.py:232-258 (vanilla VAE)."""
    d, r, n, c, _, _ = batch
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in w.items() if is_live(k)}
    ww = {**w, **leaves}
    res = forward(ww, "vae", d, r, n, c, eps_r, eps_n, True)
    terms = loss_vae(res, d, r, n, step, beta)
    l_r, l_n = glsr_regulariser(ww, res["z_r"], res["z_n"], c, d, deltas_r, deltas_n) if step > 20 else (torch.zeros(()), torch.zeros(()))
    loss = terms[0] + l_r + l_n
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    scalars = dict(loss=loss.detach(), CE_X=terms[1].detach(), CE_R=terms[2].detach(), CE_N=terms[3].detach(),
                   l_r=l_r.detach(), l_n=l_n.detach())
    return scalars, grads
