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
    if getattr(model, "precision", "f32") == "bf16":
        return greedy_decode_bf16(model, z, steps, return_logp)
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


class _DecodePlan:
    """Static buffers + the launch sequence of one (batch, steps) greedy decode on the tensor-core kernels, captured
    into a CUDA graph when the driver allows it (5 launches per step: host-side launch latency is what bounds the
    eager loop)."""

    def __init__(self, model, B, steps, return_logp):
        from .ops_bf16 import BF16, cast_bf16, tc_gemm
        from ._lib import FnGruChainBf16
        dev = model._check_device()
        D = model._dims
        H, V, G = D["H"], D["V"], D["G"]
        K3 = 3 * H
        c1, c2, lo = model.grucell_g, model.grucell_g_2, model.linear_out_g
        In1 = c1.weight_ih.shape[1]
        self.z = torch.empty((B, G), dtype=F32, device=dev)
        self.tokens = torch.empty((steps, B), dtype=torch.int32, device=dev)
        self.out = torch.empty((B, steps, V), dtype=F32, device=dev) if return_logp else None
        proj1 = torch.empty((B, K3), dtype=F32, device=dev)
        h0 = torch.empty((B, H), dtype=F32, device=dev)
        hs1 = torch.empty((steps + 1, B, H), dtype=BF16, device=dev)       # slab i = cell-1 state before step i
        hs2 = torch.empty((steps + 1, B, H), dtype=BF16, device=dev)
        tok0 = torch.full((1, B), V - 1, dtype=torch.int32, device=dev)
        dense = torch.empty((B, K3), dtype=BF16, device=dev)
        logits = torch.empty((B, V), dtype=F32, device=dev)
        logp = torch.empty((B, 1, V), dtype=F32, device=dev)
        bar = torch.empty(64, dtype=torch.uint8, device=dev)
        slab = B * H * 2
        z = self.z

        persistent = (3 * (H // 32) + (V + 95) // 96) <= 148 and H % 64 == 0 and getattr(model, "decode_persistent", True)
        if persistent:
            gi2 = torch.empty((steps, B, K3), dtype=BF16, device=dev)
            logits_all = torch.empty((steps, B, V), dtype=F32, device=dev) if return_logp else None
            nws = LIB.call("fn_decode_greedy_ws_bytes", B, steps, H, V)
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)

        def run_persistent():
            st = stream_ptr(dev)
            w1 = cast_bf16(c1.weight_hh, K3, H, H, 1)
            w2 = cast_bf16(c2.weight_hh, K3, H, H, 1)
            wi2 = cast_bf16(c2.weight_ih, K3, H, H, 1)
            wo = cast_bf16(lo.weight, V, H, H, 1)
            emb1 = cast_bf16(c1.weight_ih, V, K3, 1, In1)
            gemm(z, 0, G, 1, c1.weight_ih, V, 1, In1, proj1, 0, K3, c1.bias_ih, B, K3, G)
            gemm(z, 0, G, 1, model.linear_init_global.weight, 0, 1, G, h0, 0, H, model.linear_init_global.bias, B, H, G)
            LIB.call("fn_cast_bf16", _p(h0), H, 1, _p(hs1), H, B, H, st)
            # ONE persistent kernel for all steps (cell 1 -> cell 2 -> projection -> arg-max -> next token on the device)
            LIB.call("fn_decode_greedy_bf16", _p(w1), _p(c1.bias_hh), _p(emb1), _p(proj1), _p(wi2), _p(c2.bias_ih), _p(w2),
                     _p(c2.bias_hh), _p(wo), _p(lo.bias), _p(hs1), _p(hs2), _p(gi2), B, steps, H, V, V - 1, _p(self.tokens),
                     _p(logits_all), _p(ws), nws, st)
            if self.out is not None:
                LIB.call("fn_vocab_logsoftmax_fwd", _p(logits_all), B, steps, V, _p(self.out), st)
            self._keep = (w1, w2, wi2, wo, emb1)

        def run_stepwise():
            st = stream_ptr(dev)
            # per-call constants: bf16 weights, embedding table, time-invariant projection, initial state
            w1 = cast_bf16(c1.weight_hh, K3, H, H, 1)
            w2 = cast_bf16(c2.weight_hh, K3, H, H, 1)
            wi2 = cast_bf16(c2.weight_ih, K3, H, H, 1)
            wo = cast_bf16(lo.weight, V, H, H, 1)
            emb1 = cast_bf16(c1.weight_ih, V, K3, 1, In1)
            gemm(z, 0, G, 1, c1.weight_ih, V, 1, In1, proj1, 0, K3, c1.bias_ih, B, K3, G)
            gemm(z, 0, G, 1, model.linear_init_global.weight, 0, 1, G, h0, 0, H, model.linear_init_global.bias, B, H, G)
            LIB.call("fn_cast_bf16", _p(h0), H, 1, _p(hs1), H, B, H, st)
            ch1, ch2 = (FnGruChainBf16 * 1)(), (FnGruChainBf16 * 1)()
            ch1[0].w_hh, ch1[0].b_hh = w1.data_ptr(), c1.bias_hh.data_ptr()
            ch1[0].emb, ch1[0].proj, ch1[0].proj_ld = emb1.data_ptr(), proj1.data_ptr(), K3
            ch2[0].w_hh, ch2[0].b_hh = w2.data_ptr(), c2.bias_hh.data_ptr()
            ch2[0].dense = dense.data_ptr()
            tok = tok0
            for i in range(steps):
                ch1[0].ids, ch1[0].hsx = tok.data_ptr(), hs1.data_ptr() + i * slab
                LIB.call("fn_gru_seq_fwd_bf16", ch1, 1, B, 1, H, _p(bar), 64, st)
                tc_gemm(hs1, (i + 1) * B * H, H, 0, wi2, 0, H, 0, dense, 0, K3, c2.bias_ih, B, K3, H)
                if i == 0:
                    hs2[0].copy_(hs1[1])                               # step 0: hx[1] <- the new hx[0]
                ch2[0].hsx = hs2.data_ptr() + i * slab
                LIB.call("fn_gru_seq_fwd_bf16", ch2, 1, B, 1, H, _p(bar), 64, st)
                tc_gemm(hs2, (i + 1) * B * H, H, 0, wo, 0, H, 0, logits, 0, V, lo.bias, B, V, H)
                tok = self.tokens[i:i + 1]
                if self.out is not None:
                    LIB.call("fn_vocab_logsoftmax_fwd", _p(logits), B, 1, V, _p(logp), st)
                    LIB.call("fn_onehot_to_ids", _p(logp), B, 1, V, _p(tok), st)
                    self.out[:, i, :].copy_(logp[:, 0, :])
                else:                                                  # arg-max of the logits = arg-max of the log-probs
                    LIB.call("fn_onehot_to_ids", _p(logits), B, 1, V, _p(tok), st)
            self._keep = (w1, w2, wi2, wo, emb1)

        run = run_persistent if persistent else run_stepwise
        self.persistent = persistent
        self.run = run
        self.graph = None
        if not persistent and getattr(model, "decode_cuda_graph", True):
            try:
                run()                                                  # eager warm-up (module loading, attributes)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run()
                self.graph = g
            except Exception:                                          # capture not possible here: stay eager
                self.graph = None
                torch.cuda.synchronize(dev)

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run()


@torch.no_grad()
def greedy_decode_bf16(model, z, steps, return_logp=True):
    """Same loop on the tensor-core kernels (bf16 operands, fp32 accumulation): both cells are single-step calls of
    the tcgen05 gate block, the two projections are fn_tc_gemm_bf16.  The launch sequence of a (batch, steps) shape is
    built once and replayed as a CUDA graph.  Batches larger than 256 rows are decoded in 256-row groups (the gate
    block's chain limit)."""
    dev = model._check_device()
    z = ops._f32c(z.to(dev))
    Btot = z.shape[0]
    if Btot > 256:
        outs = [greedy_decode_bf16(model, z[i:i + 256], steps, return_logp) for i in range(0, Btot, 256)]
        lp = torch.cat([o[0] for o in outs], 0) if return_logp else None
        return lp, torch.cat([o[1] for o in outs], 0)
    cache = model.__dict__.setdefault("_decode_plans", {})
    key = (Btot, steps, bool(return_logp), model.linear_out_g.weight.data_ptr(), getattr(model, "decode_persistent", True))
    plan = cache.get(key)
    if plan is None:
        if len(cache) >= 4:
            cache.clear()
        plan = cache[key] = _DecodePlan(model, Btot, steps, return_logp)
    plan.z.copy_(z)
    plan()
    return (plan.out.clone() if return_logp else None), plan.tokens.t().contiguous().long()
