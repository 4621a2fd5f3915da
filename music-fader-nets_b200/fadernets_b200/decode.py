"""Eval-mode greedy decode of the global decoder (reference gmm_model.py:119-149 with
`_sampling` :73-80; callers test_class.py:233-254 and arousal_transfer.ipynb cells 15/17)."""
from __future__ import annotations

import torch

from . import ops
from ._lib import LIB, FnGruChain, stream_ptr
from .ops import F32, _p, gemm


@torch.no_grad()
def greedy_decode(model, z, steps, return_logp=True):
    """Per step: token gather -> cell 1 -> cell 2 -> vocabulary projection -> log-softmax ->
    first-max arg-max -> next token.  Returns (log-probs (B,steps,V) | None, tokens (B,steps))."""
    dev = model._check_device()
    D = model._dims
    H, V, G = D["H"], D["V"], D["G"]
    z = ops._f32c(z.to(dev))
    B = z.shape[0]
    st = stream_ptr(dev)
    c1, c2, lo = model.grucell_g, model.grucell_g_2, model.linear_out_g
    In1 = c1.weight_ih.shape[1]
    K3 = 3 * H

    emb1 = torch.empty((V, K3), dtype=F32, device=dev)
    LIB.call("fn_transpose_f32", _p(c1.weight_ih), In1, _p(emb1), K3, K3, V, 0, st)
    proj1 = torch.empty((B, K3), dtype=F32, device=dev)
    gemm(z, 0, G, 1, c1.weight_ih, V, 1, In1, proj1, 0, K3, c1.bias_ih, B, K3, G)
    h1 = torch.empty((B, H), dtype=F32, device=dev)
    gemm(z, 0, G, 1, model.linear_init_global.weight, 0, 1, G, h1, 0, H, model.linear_init_global.bias, B, H, G)

    tok = torch.full((1, B), V - 1, dtype=torch.int32, device=dev)
    h1n = torch.empty((B, H), dtype=F32, device=dev)
    h2 = torch.empty((B, H), dtype=F32, device=dev)
    h2n = torch.empty((B, H), dtype=F32, device=dev)
    dense = torch.empty((B, K3), dtype=F32, device=dev)
    logits = torch.empty((B, V), dtype=F32, device=dev)
    logp = torch.empty((B, 1, V), dtype=F32, device=dev)
    out = torch.empty((B, steps, V), dtype=F32, device=dev) if return_logp else None
    tokens = torch.empty((steps, B), dtype=torch.int32, device=dev)
    bar = torch.empty(64, dtype=torch.uint8, device=dev)

    ch1, ch2 = (FnGruChain * 1)(), (FnGruChain * 1)()
    ch1[0].w_hh, ch1[0].b_hh = c1.weight_hh.data_ptr(), c1.bias_hh.data_ptr()
    ch1[0].emb, ch1[0].proj, ch1[0].proj_ld = emb1.data_ptr(), proj1.data_ptr(), K3
    ch2[0].w_hh, ch2[0].b_hh = c2.weight_hh.data_ptr(), c2.bias_hh.data_ptr()
    ch2[0].dense = dense.data_ptr()
    for i in range(steps):
        ch1[0].ids, ch1[0].h0, ch1[0].hs = tok.data_ptr(), h1.data_ptr(), h1n.data_ptr()
        LIB.call("fn_gru_seq_fwd_f32", ch1, 1, B, 1, H, _p(bar), 64, st)
        gemm(h1n, 0, H, 1, c2.weight_ih, 0, 1, H, dense, 0, K3, c2.bias_ih, B, K3, H)
        ch2[0].h0 = (h1n if i == 0 else h2).data_ptr()           # step 0: hx[1] <- the new hx[0]
        ch2[0].hs = h2n.data_ptr()
        LIB.call("fn_gru_seq_fwd_f32", ch2, 1, B, 1, H, _p(bar), 64, st)
        gemm(h2n, 0, H, 1, lo.weight, 0, 1, H, logits, 0, V, lo.bias, B, V, H)
        LIB.call("fn_vocab_logsoftmax_fwd", _p(logits), B, 1, V, _p(logp), st)
        tok = tokens[i:i + 1]
        LIB.call("fn_onehot_to_ids", _p(logp), B, 1, V, _p(tok), st)
        if out is not None:
            out[:, i, :].copy_(logp[:, 0, :])
        h1, h1n = h1n, h1
        h2, h2n = h2n, h2
    return out, tokens.t().contiguous().long()
