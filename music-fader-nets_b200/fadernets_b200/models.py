"""Host-side mirror of the reference model classes for the GM-VAE / VAE hot path.

`MusicAttrRegGMVAE` (reference gmm_model.py:10-259) and `MusicAttrRegVAE` (model_v2.py:9-171)
keep the reference's constructor, attributes, method names, return tuples and state_dict keys
(so params/*.pt load strictly), but every tensor operation of forward / encode / sub_decoders /
global_decoder runs in the sm_100a CUDA library through fadernets_b200.ops.  torch.nn modules
are used ONLY as parameter containers (same registration order => same default init under a
seed => same state_dict); their forward() is never called.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn
from torch.distributions import Normal

from . import ops
from ._lib import FaderNetsError, require_cuda
from .ops import ChainSpec, GruGroupFn, LatentHeadFn, QyXFn, TimeLogSoftmaxFn, VocabLogSoftmaxFn, linear
from .ops_bf16 import DecoderStackBf16Fn, GruGroupBf16Fn, linear_bf16, GruGroupBf16
from .ops_x3 import DecoderStackX3Fn, GruGroupX3, linear_x3

PRECISIONS = ("f32", "bf16", "bf16x3")

START_TOKEN_FROM_END = 1          # decoder start symbol = one-hot of the LAST vocabulary index (gmm_model.py:120-121)
CDTL_DIMS = 24                    # chroma conditioning width hard-coded by the reference (gmm_model.py:53)

# parameters that exist only for checkpoint compatibility (never used by forward)
DEAD_PREFIXES = ("gru_c.", "gru_d_c.", "c_r.", "c_n.", "mu_c.", "var_c.", "linear_init_c.", "linear_out_c.")


class _FaderBase(nn.Module):
    variant = "base"

    # ---------------------------------------------------------------- construction
    def _register_reference_parameters(self, roll, rhythm, note, chroma, H, Z, classifiers_first: bool):
        """Registers sub-modules in the reference's order (parameter containers only)."""
        def bi(inp):
            return nn.GRU(inp, H, batch_first=True, bidirectional=True)

        def uni(inp):
            return nn.GRU(inp, H, batch_first=True)

        plan = [("gru_r", lambda: bi(roll)), ("gru_n", lambda: bi(roll)), ("gru_c", lambda: bi(roll))]
        cls = [("c_r", lambda: nn.Linear(Z, 3)), ("c_n", lambda: nn.Linear(Z, 3))]
        subdec = [("gru_d_r", lambda: uni(Z + rhythm)), ("gru_d_n", lambda: uni(Z + note)),
                  ("gru_d_c", lambda: uni(Z + chroma))]
        plan += (cls + subdec) if classifiers_first else (subdec + cls)
        for a in "rnc":
            plan += [(f"mu_{a}", lambda: nn.Linear(2 * H, Z)), (f"var_{a}", lambda: nn.Linear(2 * H, Z))]
        G = 2 * Z + CDTL_DIMS
        plan += [("linear_init_global", lambda: nn.Linear(G, H)), ("grucell_g", lambda: nn.GRUCell(G + roll, H)),
                 ("grucell_g_2", lambda: nn.GRUCell(H, H))]
        plan += [(f"linear_init_{a}", lambda: nn.Linear(Z, H)) for a in "rnc"]
        plan += [("linear_out_r", lambda: nn.Linear(H, rhythm)), ("linear_out_n", lambda: nn.Linear(H, note)),
                 ("linear_out_c", lambda: nn.Linear(Z, chroma)), ("linear_out_g", lambda: nn.Linear(H, roll))]
        for name, make in plan:
            setattr(self, name, make())
        self._dims = dict(V=roll, R=rhythm, N=note, C=chroma, H=H, Z=Z, G=G)
        self._flat = None
        self._flat_grad = None
        self.host_rng = True          # draw eps exactly like the reference (CPU default generator)
        self.precision = "f32"        # "f32": exact-parity SIMT path; "bf16": tcgen05 tensor-core path

    # ---------------------------------------------------------------- flat parameter storage
    def live_parameters(self):
        """Parameters that receive gradients, in flat-buffer order: decoder-side first, the encoders last, so that the
        data-parallel gradient buckets are contiguous ranges (parallel.decoder_first)."""
        from .parallel import decoder_first
        return decoder_first([(n, p) for n, p in self.named_parameters() if p.requires_grad and not n.startswith(DEAD_PREFIXES)])

    def _flat_ok(self):
        if self._flat is None:
            return False
        off = 0
        base = self._flat.data_ptr()
        for _, p in self.live_parameters():              # every live parameter must still be a view of the flat buffer
            if p.data_ptr() != base + off * 4 or p.device != self._flat.device:
                return False
            off += ((p.numel() + 3) // 4) * 4
        return True

    def flatten_parameters_(self):
        """Packs the live parameters (and their grads) into two flat fp32 buffers so the
        clip+Adam kernel and the gradient all-reduce see one contiguous range.  Idempotent."""
        if self._flat_ok():
            return self._flat, self._flat_grad
        live = self.live_parameters()
        dev = live[0][1].device
        total = sum(((p.numel() + 3) // 4) * 4 for _, p in live)
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for _, p in live:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                p.grad = grad[off:off + n].view(p.shape)
                off += ((n + 3) // 4) * 4
        self._flat, self._flat_grad = flat, grad
        return flat, grad

    def zero_grad_flat(self):
        """optimizer.zero_grad() for the flat layout: one memset, grads stay attached as views."""
        flat, grad = self.flatten_parameters_()
        grad.zero_()
        off = 0
        for _, p in self.live_parameters():
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != grad.data_ptr() + off * 4:
                p.grad = grad[off:off + n].view(p.shape)
            off += ((n + 3) // 4) * 4

    def set_precision(self, precision: str):
        """"f32" (default): every product in fp32 FMA -- the 1e-3 parity mode.  "bf16": the T-scale products
        (GRU gate block, vocabulary / attribute heads, their gradients) on the tcgen05 tensor cores with bf16
        operands and fp32 accumulation (BASELINE configs 3-5); parameters and reductions stay fp32.  "bf16x3": the same
        tensor-core kernels with every fp32 operand carried as two bf16 planes (hi, lo) and three plane products per
        product -- fp32-level results (meets the 1e-3 bar of "f32") at tensor-core speed; eval-mode greedy decode runs
        the exact fp32 path."""
        if precision not in PRECISIONS:
            raise FaderNetsError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        return self

    def _gru(self):
        return GruGroupBf16 if self.precision == "bf16" else GruGroupX3 if self.precision == "bf16x3" else GruGroupFn

    def _tlinear(self, x, lin):
        """Linear over a T*B-row activation (time-major hidden states)."""
        if self.precision == "bf16":
            return linear_bf16(x, lin.weight, lin.bias)
        if self.precision == "bf16x3":
            return linear_x3(x, lin.weight, lin.bias)
        return linear(x, lin.weight, lin.bias)

    # ---------------------------------------------------------------- helpers
    def _check_device(self):
        p = self.linear_out_g.weight
        if not p.is_cuda:
            raise FaderNetsError("fadernets_b200 models run on CUDA (B200) only; call model.cuda() first. "
                                 "There is no CPU fallback.")
        self.flatten_parameters_()
        return p.device

    def _token_buffers(self, x: torch.Tensor):
        """x: one-hot (B,T,V) fp32 or ids (B,T) int64 -> int32 [T+1,B] buffer whose row 0 is the
        decoder start token; rows 1.. are the tokens (time-major)."""
        dev = self._check_device()
        V = self._dims["V"]
        require_cuda(x)
        if x.dim() == 3:
            B, T, _ = x.shape
            buf = torch.empty((T + 1, B), dtype=torch.int32, device=dev)
            ops.LIB.call("fn_onehot_to_ids", ops._p(ops._f32c(x)), B, T, V, ops._p(buf, B), ops.stream_ptr(dev))
        else:
            B, T = x.shape
            buf = torch.empty((T + 1, B), dtype=torch.int32, device=dev)
            xi = x.long().contiguous()
            ops.LIB.call("fn_ids_to_time_major", ops._p(xi), B, T, 0, 0, ops._p(buf, B), ops.stream_ptr(dev))
            ops.clamp_index(buf[1:], V)                 # nn.Embedding / one-hot scatter of the reference raise on a bad id
        buf[0].fill_(V - START_TOKEN_FROM_END)
        return buf

    def _attr_ids(self, a: torch.Tensor, dims: int):
        if a.dim() == 3:
            return ops.onehot_to_ids_tm(a)
        return ops.ids_to_tm(a, dims=dims)

    def _draw_eps(self, B, Z, dev):
        """Noise for the reparameterisation (gmm_model.py:229-235): the reference samples on the CPU
        default generator and copies to the device; reproduced so a seeded run matches draw for draw."""
        if self.host_rng:
            e = Normal(0, 1).sample(sample_shape=torch.Size((B, Z)))
            return e.to(dev, non_blocking=True)
        return torch.randn((B, Z), device=dev)

    # ---------------------------------------------------------------- encoder
    def _encoder_heads(self, ids_tm: torch.Tensor):
        T, B = ids_tm.shape
        H, V = self._dims["H"], self._dims["V"]
        specs, tensors = [], []
        for gi, g in enumerate((self.gru_r, self.gru_n)):
            for sfx, rev in (("", False), ("_reverse", True)):
                specs.append(ChainSpec(emb_cols=(0, V), ids=ids_tm, reverse=rev, final=(gi, H if rev else 0)))
                tensors += [getattr(g, f"weight_ih_l0{sfx}"), getattr(g, f"bias_ih_l0{sfx}"),
                            getattr(g, f"weight_hh_l0{sfx}"), getattr(g, f"bias_hh_l0{sfx}")]
        hcat_r, hcat_n = self._gru().apply(specs, B, T, H, (2 * H, 2 * H), *tensors)
        mu_r, pre_r = linear(hcat_r, self.mu_r.weight, self.mu_r.bias), linear(hcat_r, self.var_r.weight, self.var_r.bias)
        mu_n, pre_n = linear(hcat_n, self.mu_n.weight, self.mu_n.bias), linear(hcat_n, self.var_n.weight, self.var_n.bias)
        return mu_r, pre_r, mu_n, pre_n

    def _encode_dists(self, x):
        buf = self._token_buffers(x)
        mu_r, pre_r, mu_n, pre_n = self._encoder_heads(buf[1:])
        s_r, s_n = ops.ExpFn.apply(pre_r), ops.ExpFn.apply(pre_n)
        return Normal(mu_r, s_r, validate_args=False), Normal(mu_n, s_n, validate_args=False)

    # ---------------------------------------------------------------- decoders
    def _decoder_states(self, r_tm, z_r, n_tm, z_n, d1_tm, zc):
        """One persistent launch for the two sub-decoder GRUs and global cell 1 (teacher-forced),
        a second one for global cell 2.  Any of the three chains may be omitted (None ids)."""
        D = self._dims
        H, V, Z, G = D["H"], D["V"], D["Z"], D["G"]
        specs, tensors, names = [], [], []
        if r_tm is not None:
            T, B = r_tm.shape
            specs.append(ChainSpec(emb_cols=(0, D["R"]), ids=r_tm, z_cols=(D["R"], Z), h0="tensor", want_hs=True))
            g = self.gru_d_r
            tensors += [g.weight_ih_l0, g.bias_ih_l0, g.weight_hh_l0, g.bias_hh_l0, z_r,
                        linear(z_r, self.linear_init_r.weight, self.linear_init_r.bias)]
            names.append("r")
        if n_tm is not None:
            T, B = n_tm.shape
            specs.append(ChainSpec(emb_cols=(0, D["N"]), ids=n_tm, z_cols=(D["N"], Z), h0="tensor", want_hs=True))
            g = self.gru_d_n
            tensors += [g.weight_ih_l0, g.bias_ih_l0, g.weight_hh_l0, g.bias_hh_l0, z_n,
                        linear(z_n, self.linear_init_n.weight, self.linear_init_n.bias)]
            names.append("n")
        if d1_tm is not None:
            T, B = d1_tm.shape
            specs.append(ChainSpec(emb_cols=(0, V), ids=d1_tm, z_cols=(V, G), h0="tensor", want_hs=True))
            c = self.grucell_g
            tensors += [c.weight_ih, c.bias_ih, c.weight_hh, c.bias_hh, zc,
                        linear(zc, self.linear_init_global.weight, self.linear_init_global.bias)]
            names.append("g")
        if self.precision in ("bf16", "bf16x3") and names == ["r", "n", "g"] and B <= 256:
            # the whole decoder stack as one wavefront (cell 2 one time segment behind cell 1 in the same launches)
            c2 = self.grucell_g_2
            stack = DecoderStackBf16Fn if self.precision == "bf16" else DecoderStackX3Fn
            hs_r, hs_n, hs_g2 = stack.apply(tuple(specs), B, T, H, *tensors, c2.weight_ih, c2.bias_ih,
                                            c2.weight_hh, c2.bias_hh)
            return {"r": hs_r, "n": hs_n, "g2": hs_g2}
        hs = dict(zip(names, self._gru().apply(specs, B, T, H, (), *tensors)))
        if "g" in hs:
            c2 = self.grucell_g_2
            (hs["g2"],) = self._gru().apply([ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)], B, T, H, (),
                                            c2.weight_ih, c2.bias_ih, c2.weight_hh, c2.bias_hh, hs["g"])
        return hs

    def _sub_decoder_outputs(self, hs):
        lr = self._tlinear(hs["r"], self.linear_out_r)
        ln = self._tlinear(hs["n"], self.linear_out_n)
        return TimeLogSoftmaxFn.apply(lr), TimeLogSoftmaxFn.apply(ln)

    def _global_logits(self, hs):
        return self._tlinear(hs["g2"], self.linear_out_g)       # [T,B,V]

    def sub_decoders(self, rhythm, z_r, note, z_n):
        """gmm_model.py:100-117 / model_v2.py:99-116 (log-softmax over the time axis)."""
        hs = self._decoder_states(self._attr_ids(rhythm, self._dims["R"]), z_r,
                                  self._attr_ids(note, self._dims["N"]), z_n, None, None)
        r_out, n_out = self._sub_decoder_outputs(hs)
        return (r_out, n_out, 0, 0) if self.variant == "gmvae" else (r_out, n_out)

    def _sampling(self, x):
        """One-hot of the first arg-max per row (gmm_model.py:73-80)."""
        require_cuda(x)
        ids = ops.onehot_to_ids_tm(x.unsqueeze(1)).view(-1).long()
        return ops.ids_to_onehot(ids, x.shape[1])

    def global_decoder(self, z, steps):
        """gmm_model.py:119-149 / model_v2.py:118-143.  training mode: teacher forcing from
        self.sample (the reference's coin `p < eps=100` always succeeds); eval: greedy arg-max."""
        dev = self._check_device()
        require_cuda(z)
        if self.training:
            if self.sample is None:
                raise FaderNetsError("global_decoder in training mode needs self.sample (teacher tokens)")
            if self.host_rng:
                torch.rand(steps)                                 # the reference's per-step coin (:140)
            buf = self._sample_tokens if getattr(self, "_sample_tokens", None) is not None and \
                self._sample_src is self.sample else self._token_buffers(self.sample)
            if buf.shape[0] - 1 < steps:
                raise FaderNetsError(f"teacher sequence shorter ({buf.shape[0] - 1}) than steps ({steps})")
            hs = self._decoder_states(None, None, None, None, buf[:steps], z)
            return VocabLogSoftmaxFn.apply(self._global_logits(hs))
        out, _ = self.decode_greedy(z, steps)
        return out

    @torch.no_grad()
    def decode_greedy(self, z, steps, return_logp=True):
        """Eval-mode decode: returns (log-probs (B,steps,V) or None, tokens (B,steps) int64)."""
        from .decode import greedy_decode
        return greedy_decode(self, z, steps, return_logp)

    # ---------------------------------------------------------------- full forward
    def _forward_common(self, x, rhythm, note, chroma):
        dev = self._check_device()
        D = self._dims
        buf = self._token_buffers(x)
        if self.training:
            self.sample = x
            self._sample_tokens, self._sample_src = buf, x
        T, B = buf.shape[0] - 1, buf.shape[1]
        mu_r, pre_r, mu_n, pre_n = self._encoder_heads(buf[1:])
        eps_r = self._draw_eps(B, D["Z"], dev)
        eps_n = self._draw_eps(B, D["Z"], dev)
        s_r, z_r = LatentHeadFn.apply(mu_r, pre_r, eps_r)
        s_n, z_n = LatentHeadFn.apply(mu_n, pre_n, eps_n)
        dis_r, dis_n = Normal(mu_r, s_r, validate_args=False), Normal(mu_n, s_n, validate_args=False)
        zc = torch.cat([z_r, z_n, chroma.to(dev).float()], dim=1)
        r_tm, n_tm = self._attr_ids(rhythm, D["R"]), self._attr_ids(note, D["N"])
        if self.training:
            if self.host_rng:
                torch.rand(T)                                     # decoder coin flips (gmm_model.py:140)
            hs = self._decoder_states(r_tm, z_r, n_tm, z_n, buf[:T], zc)
            out = VocabLogSoftmaxFn.apply(self._global_logits(hs))
        else:
            hs = self._decoder_states(r_tm, z_r, n_tm, z_n, None, None)
            out = self.global_decoder(zc, T)
        r_out, n_out = self._sub_decoder_outputs(hs)
        self._last_logits = None
        return out, r_out, n_out, dis_r, dis_n, z_r, z_n


class MusicAttrRegGMVAE(_FaderBase):
    """Drop-in for reference gmm_model.py:10-259 (GM-VAE with a K-component mixture prior)."""
    variant = "gmvae"

    def __init__(self, roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims, n_step, n_component=4):
        super().__init__()
        self.n_component = n_component
        self.latent_dim = z_dims
        self.roll_dims = roll_dims
        self.eps = 100
        self.sample = None
        self._register_reference_parameters(roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims,
                                            classifiers_first=True)
        self._build_mu_lookup()
        self._build_logvar_lookup(pow_exp=-2)

    def _build_mu_lookup(self):
        for name in ("mu_r_lookup", "mu_n_lookup"):            # xavier-uniform means (gmm_model.py:151-165)
            emb = nn.Embedding(self.n_component, self.latent_dim)
            nn.init.xavier_uniform_(emb.weight)
            setattr(self, name, emb)

    def _build_logvar_lookup(self, pow_exp=0, logvar_trainable=False):
        for name in ("logvar_r_lookup", "logvar_n_lookup"):    # constant log sigma^2 (gmm_model.py:167-183)
            emb = nn.Embedding(self.n_component, self.latent_dim)
            nn.init.constant_(emb.weight, math.log(math.exp(pow_exp) ** 2))
            emb.weight.requires_grad = logvar_trainable
            setattr(self, name, emb)

    def encode(self, x):
        """gmm_model.py:82-98: (Normal(mu_r, exp(.)), Normal(mu_n, exp(.)))."""
        return self._encode_dists(x)

    def approx_qy_x(self, z, mu_lookup, logvar_lookup, n_component):
        """gmm_model.py:194-218 -> (logLogit_qy_x, qy_x)."""
        ll, qy, _ = QyXFn.apply(z, mu_lookup.weight[:n_component], logvar_lookup.weight[:n_component])
        return ll, qy

    def forward(self, x, rhythm, note, chroma):
        """gmm_model.py:220-259; x may also be int64 token ids (B,T) (then rhythm/note are ids too)."""
        out, r_out, n_out, dis_r, dis_n, z_r, z_n = self._forward_common(x, rhythm, note, chroma)
        ll_r, qy_r, y_r = QyXFn.apply(z_r, self.mu_r_lookup.weight, self.logvar_r_lookup.weight)
        ll_n, qy_n, y_n = QyXFn.apply(z_n, self.mu_n_lookup.weight, self.logvar_n_lookup.weight)
        return ((out, r_out, n_out, 0, 0), (dis_r, dis_n), (z_r, z_n), (ll_r, ll_n), (qy_r, qy_n), (y_r, y_n))


class MusicAttrRegVAE(_FaderBase):
    """Drop-in for reference model_v2.py:9-171 (vanilla VAE, standard-normal prior)."""
    variant = "vae"

    def __init__(self, roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims, n_step, k=1000):
        super().__init__()
        self._register_reference_parameters(roll_dims, rhythm_dims, note_dims, chroma_dims, hidden_dims, z_dims,
                                            classifiers_first=False)
        self.n_step = n_step
        self.roll_dims = roll_dims
        self.hidden_dims = hidden_dims
        self.eps = 100
        self.rhythm_dims = rhythm_dims
        self.sample = None
        self.iteration = 0
        self.z_dims = z_dims
        self.k = torch.FloatTensor([k])

    def encoder(self, x):
        """model_v2.py:81-97."""
        return self._encode_dists(x)

    def forward(self, x, rhythm, note, chroma):
        """model_v2.py:145-171."""
        if self.training:
            self.iteration += 1
        out, r_out, n_out, dis_r, dis_n, z_r, z_n = self._forward_common(x, rhythm, note, chroma)
        return ((out, r_out, n_out), (dis_r, dis_n), (z_r, z_n))
