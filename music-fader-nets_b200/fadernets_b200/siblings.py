"""Sibling models of the reference on the same CUDA blocks (reference model_v2.py:174-586):

  MusicAttrSingleVAE  (:174-285)  one bidirectional encoder, 2Z latent, Pati et al. regulariser on z[:,0] / z[:,1]
  MusicAttrCVAE       (:288-423)  encoder input = event one-hot + (r_density, n_density); decoder conditioned on them
  MusicAttrFaderNets  (:438-586)  CVAE decoder + adversarial discriminators behind a gradient reversal

Same constructor arguments, attributes, method names, return tuples and state_dict keys as the reference classes.  They
share the encoder (GruGroupFn: token-embedding gather + time-invariant projection of the appended columns), the latent
head and the two-cell global decoder (teacher-forced chains / one-kernel greedy decode) with MusicAttrRegGMVAE; the only
new device code is three element-wise kernels (relu x dropout mask, gradient scale, batch MSE).  Draw order of the CPU
generator follows the reference: repar noise, (FaderNets: dropout mask r, dropout mask n), then T coin flips."""
from __future__ import annotations

import torch
from torch import nn
from torch.distributions import Normal

from . import ops
from .models import _FaderBase
from .ops import ChainSpec, GradReverseFn, LatentHeadFn, ReluMaskFn, VocabLogSoftmaxFn, linear


class _SiblingBase(_FaderBase):
    ENC = "gru_e"            # name of the encoder module
    ZW = 1                   # latent width in units of z_dims
    CDTL = 2                 # conditioning columns appended to z for the decoder
    ENC_EXTRA = 0            # dense columns appended to every encoder input step

    def _register(self, roll, H, Z, extra_modules):
        self._dims = dict(V=roll, R=3, N=16, C=24, H=H, Z=Z * self.ZW, G=Z * self.ZW + self.CDTL)
        setattr(self, self.ENC, nn.GRU(roll + self.ENC_EXTRA, H, batch_first=True, bidirectional=True))
        for name, make in extra_modules:
            setattr(self, name, make())
        G = self._dims["G"]
        self.linear_init_global = nn.Linear(G, H)
        self.grucell_g = nn.GRUCell(G + roll, H)
        self.grucell_g_2 = nn.GRUCell(H, H)
        self.linear_out_g = nn.Linear(H, roll)
        self._flat = self._flat_grad = None
        self.host_rng = True
        self.precision = "f32"

    def _common_attrs(self, roll_dims, rhythm_dims, hidden_dims, z_dims, n_step, k):
        self.n_step, self.roll_dims, self.hidden_dims = n_step, roll_dims, hidden_dims
        self.eps, self.rhythm_dims, self.sample, self.iteration, self.z_dims = 100, rhythm_dims, None, 0, z_dims
        self.k = torch.FloatTensor([k])

    # -- encoder: one bidirectional GRU -> [h_fwd(T-1) | h_bwd(0)] -> mu, exp(var)
    def _encode(self, ids_tm, extra=None):
        T, B = ids_tm.shape
        H, V = self._dims["H"], self._dims["V"]
        g = getattr(self, self.ENC)
        specs, tensors = [], []
        for sfx, rev in (("", False), ("_reverse", True)):
            specs.append(ChainSpec(emb_cols=(0, V), ids=ids_tm, reverse=rev, final=(0, H if rev else 0),
                                   z_cols=None if extra is None else (V, extra.shape[1])))
            tensors += [getattr(g, f"weight_ih_l0{sfx}"), getattr(g, f"bias_ih_l0{sfx}"),
                        getattr(g, f"weight_hh_l0{sfx}"), getattr(g, f"bias_hh_l0{sfx}")]
            if extra is not None:
                tensors.append(extra)
        (hcat,) = self._gru().apply(specs, B, T, H, (2 * H,), *tensors)
        return linear(hcat, self.mu.weight, self.mu.bias), linear(hcat, self.var.weight, self.var.bias)

    def _decode_train_or_eval(self, buf, zc, T):
        if self.training:
            if self.host_rng:
                torch.rand(T)                                     # the reference's per-step coin (model_v2.py:251)
            hs = self._decoder_states(None, None, None, None, buf[:T], zc)
            return VocabLogSoftmaxFn.apply(self._global_logits(hs))
        return self.global_decoder(zc, T)

    def _begin(self, x):
        dev = self._check_device()
        buf = self._token_buffers(x)
        if self.training:
            self.sample = x
            self._sample_tokens, self._sample_src = buf, x
            self.iteration += 1
        return dev, buf, buf.shape[0] - 1, buf.shape[1]


class MusicAttrSingleVAE(_SiblingBase):
    """Drop-in for reference model_v2.py:174-285."""
    variant, ENC, ZW, CDTL = "singlevae", "gru", 2, 24

    def __init__(self, roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims, n_step, k=1000):
        super().__init__()
        self._register(roll_dims, hidden_dims, z_dims,
                       [("e_dropout", lambda: nn.Dropout(p=0.3)),
                        ("mu", lambda: nn.Linear(hidden_dims * 2, z_dims * 2)), ("var", lambda: nn.Linear(hidden_dims * 2, z_dims * 2))])
        self._common_attrs(roll_dims, rhythm_dims, hidden_dims, z_dims, n_step, k)

    def encoder(self, x):
        mu, pre = self._encode(self._token_buffers(x)[1:])
        return Normal(mu, ops.ExpFn.apply(pre), validate_args=False)

    def forward(self, x, chroma):
        dev, buf, T, B = self._begin(x)
        mu, pre = self._encode(buf[1:])
        s, z = LatentHeadFn.apply(mu, pre, self._draw_eps(B, self._dims["Z"], dev))
        zc = torch.cat([z, chroma.to(dev).float()], dim=1)
        return self._decode_train_or_eval(buf, zc, T), Normal(mu, s, validate_args=False), zc


class MusicAttrCVAE(_SiblingBase):
    """Drop-in for reference model_v2.py:288-423 (c_r / c_n are registered but unused, like in the reference)."""
    variant, ENC_EXTRA = "cvae", 2

    def __init__(self, roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims, n_step, k=1000):
        super().__init__()
        self._register(roll_dims, hidden_dims, z_dims,
                       [("c_r", lambda: nn.Linear(z_dims, 3)), ("c_n", lambda: nn.Linear(z_dims, 3)),
                        ("mu", lambda: nn.Linear(hidden_dims * 2, z_dims)), ("var", lambda: nn.Linear(hidden_dims * 2, z_dims))])
        self._common_attrs(roll_dims, rhythm_dims, hidden_dims, z_dims, n_step, k)

    def _cond(self, r_density, n_density, dev):
        return torch.cat([r_density.to(dev).float().reshape(-1, 1), n_density.to(dev).float().reshape(-1, 1)], dim=1)

    def encoder(self, x, r_density, n_density, chroma):
        dev = self._check_device()
        mu, pre = self._encode(self._token_buffers(x)[1:], self._cond(r_density, n_density, dev))
        return Normal(mu, ops.ExpFn.apply(pre), validate_args=False)

    def forward(self, x, rhythm, note, chroma, r_density, n_density):
        dev, buf, T, B = self._begin(x)
        cond = self._cond(r_density, n_density, dev)
        mu, pre = self._encode(buf[1:], cond)
        s, z = LatentHeadFn.apply(mu, pre, self._draw_eps(B, self._dims["Z"], dev))
        zc = torch.cat([z, cond], dim=-1)
        return self._decode_train_or_eval(buf, zc, T), Normal(mu, s, validate_args=False), zc


class MusicAttrFaderNets(_SiblingBase):
    """Drop-in for reference model_v2.py:438-586: CVAE decoder + two discriminators behind a gradient reversal."""
    variant = "fader"

    def __init__(self, roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims, n_step, k=1000):
        super().__init__()
        self._register(roll_dims, hidden_dims, z_dims,
                       [("c_r", lambda: nn.Linear(z_dims, 3)), ("c_n", lambda: nn.Linear(z_dims, 3)),
                        ("mu", lambda: nn.Linear(hidden_dims * 2, z_dims)), ("var", lambda: nn.Linear(hidden_dims * 2, z_dims)),
                        ("discriminator_r", lambda: nn.Linear(z_dims, 1)), ("discriminator_n", lambda: nn.Linear(z_dims, 1)),
                        ("dropout", lambda: nn.Dropout(p=0.3))])
        self._common_attrs(roll_dims, rhythm_dims, hidden_dims, z_dims, n_step, k)

    def encoder(self, x):
        mu, pre = self._encode(self._token_buffers(x)[1:])
        return Normal(mu, ops.ExpFn.apply(pre), validate_args=False)

    def _dropout_mask(self, B, dev):
        """Keep mask / (1 - p) of nn.Dropout(p=0.3) on a (B,1) tensor; drawn on the CPU default generator like the
        reference's CPU dropout when host_rng (so a seeded run replays the reference), on the device otherwise."""
        if not self.training:
            return None
        ones = torch.ones(B, 1) if self.host_rng else torch.ones(B, 1, device=dev)
        return torch.nn.functional.dropout(ones, 0.3, True).to(dev)

    def forward(self, x, rhythm, note, chroma, r_density, n_density):
        dev, buf, T, B = self._begin(x)
        mu, pre = self._encode(buf[1:])
        s, z = LatentHeadFn.apply(mu, pre, self._draw_eps(B, self._dims["Z"], dev))
        r_z = GradReverseFn.apply(z)
        r_out = ReluMaskFn.apply(linear(r_z, self.discriminator_r.weight, self.discriminator_r.bias), self._dropout_mask(B, dev))
        n_out = ReluMaskFn.apply(linear(r_z, self.discriminator_n.weight, self.discriminator_n.bias), self._dropout_mask(B, dev))
        zc = torch.cat([z, r_density.to(dev).float().reshape(-1, 1), n_density.to(dev).float().reshape(-1, 1)], dim=-1)
        out = self._decode_train_or_eval(buf, zc, T)
        return (out, r_out, n_out), Normal(mu, s, validate_args=False), zc
