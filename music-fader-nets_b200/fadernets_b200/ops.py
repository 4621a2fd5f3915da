"""torch.autograd glue over the C ABI (include/fadernets_b200.h).

Every Function here is a thin host-side wrapper: it allocates device buffers with torch,
hands raw pointers to the CUDA library and wires the result into autograd.  No arithmetic
of the hot path is done by torch operators.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from ._lib import LIB, FnGruChain, require_cuda, stream_ptr

F32 = torch.float32


def _p(t: Optional[torch.Tensor], off_elems: int = 0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + off_elems * t.element_size())


def _st(t: torch.Tensor):
    return stream_ptr(t.device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != F32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def gemm(A, a_off, sam, sak, B, b_off, sbk, sbn, Cm, c_off, ldc, bias, M, N, K, accumulate=False):
    # few 64x64 output tiles and a long K (latent heads, z-projection gradients): the K loop of a handful of CTAs is a latency
    # chain -- split it over enough CTAs to fill the machine (deterministic fixed-order reduction of the partials)
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    if K >= 512 and 0 < tiles <= 48:
        splits = max(1, min(K // 128, 148 // tiles))
        if splits > 1:
            nb = LIB.call("fn_gemm_f32_splitk_ws_bytes", M, N, splits)
            ws = torch.empty(nb, dtype=torch.uint8, device=Cm.device)
            LIB.call("fn_gemm_f32_splitk", _p(A, a_off), sam, sak, _p(B, b_off), sbk, sbn, _p(Cm, c_off), ldc, _p(bias),
                     M, N, K, 1 if accumulate else 0, splits, _p(ws), nb, _st(Cm))
            return
    LIB.call("fn_gemm_f32", _p(A, a_off), sam, sak, _p(B, b_off), sbk, sbn, _p(Cm, c_off), ldc, _p(bias),
             M, N, K, 1 if accumulate else 0, _st(Cm))


def col_sum(x: torch.Tensor, ld: int, rows: int, cols: int, out: torch.Tensor, accumulate=False):
    nbytes = LIB.call("fn_col_sum_scratch_bytes", rows, cols)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    LIB.call("fn_col_sum_f32", _p(x), ld, rows, cols, _p(out), 1 if accumulate else 0, _p(scratch), nbytes, _st(x))


def _reduce_scratch(dev):
    n = LIB.call("fn_reduce_scratch_bytes", 0)
    return torch.empty(n, dtype=torch.uint8, device=dev), n


# ------------------------------------------------------------------------------------------------
# index validation
# ------------------------------------------------------------------------------------------------
# The reference's F.nll_loss / nn.Embedding raise on an out-of-range index.  Here every consuming kernel clamps (no
# out-of-bounds access), offending entries are COUNTED on the device, and the mirror raises IndexError at its next
# synchronisation point (train / evaluate read their results back there; `raise_on_bad_index()` for other callers).
_BAD_INDEX = {}


def _bad_counter(dev) -> torch.Tensor:
    key = torch.device(dev).index or 0
    t = _BAD_INDEX.get(key)
    if t is None:
        t = _BAD_INDEX[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return t


def check_index(idx: torch.Tensor, hi: int):
    """Count entries of the int64 tensor `idx` outside [0, hi) (targets / labels; consumers clamp)."""
    idx = idx if idx.is_contiguous() else idx.contiguous()
    LIB.call("fn_check_index_i64", _p(idx), idx.numel(), hi, _p(_bad_counter(idx.device)), _st(idx))


def clamp_index(idx: torch.Tensor, hi: int):
    """Clamp the int32 id buffer `idx` into [0, hi) in place and count what had to be clamped."""
    LIB.call("fn_clamp_index_i32", _p(idx), idx.numel(), hi, _p(_bad_counter(idx.device)), _st(idx))


def raise_on_bad_index(dev=None):
    """Synchronises; raises IndexError if any index checked since the last call was out of range."""
    for key, t in list(_BAD_INDEX.items()):
        if dev is not None and (torch.device(dev).index or 0) != key:
            continue
        n = int(t.item())
        if n:
            t.zero_()
            raise IndexError(f"fadernets_b200: {n} token / target / label index(es) out of range "
                             "(the reference's nll_loss / Embedding raise here too)")


# ------------------------------------------------------------------------------------------------
# token plumbing (no autograd)
# ------------------------------------------------------------------------------------------------
def onehot_to_ids_tm(onehot: torch.Tensor) -> torch.Tensor:
    """(B,T,V) fp32 one-hot -> int32 [T,B] first-max ids."""
    require_cuda(onehot)
    onehot = _f32c(onehot)
    B, T, V = onehot.shape
    out = torch.empty((T, B), dtype=torch.int32, device=onehot.device)
    LIB.call("fn_onehot_to_ids", _p(onehot), B, T, V, _p(out), _st(out))
    return out


def ids_to_onehot(ids: torch.Tensor, dims: int) -> torch.Tensor:
    require_cuda(ids)
    ids = ids.long().contiguous()
    if ids.dim() == 1:
        B, T = ids.shape[0], 1
    else:
        B, T = ids.shape
    out = torch.empty((B, T, dims), dtype=F32, device=ids.device)
    LIB.call("fn_ids_to_onehot", _p(ids), B, T, dims, _p(out), _st(out))
    return out if ids.dim() > 1 else out.view(B, dims)


def ids_to_tm(ids: torch.Tensor, shift: int = 0, start_token: int = 0, dims: Optional[int] = None) -> torch.Tensor:
    """int64 (B,T) -> int32 [T,B]; shift=1 prepends the start token (teacher-forced decoder input).
    `dims`: vocabulary size -- ids are validated (clamped + counted, see check_index)."""
    require_cuda(ids)
    ids = ids.long().contiguous()
    B, T = ids.shape
    out = torch.empty((T, B), dtype=torch.int32, device=ids.device)
    LIB.call("fn_ids_to_time_major", _p(ids), B, T, shift, start_token, _p(out), _st(out))
    if dims is not None:
        clamp_index(out, dims)
    return out


# ------------------------------------------------------------------------------------------------
# Linear
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b for x [..., K] (contiguous), W [N,K], b [N]."""

    @staticmethod
    def forward(ctx, x, w, b):
        require_cuda(x, w)
        x = _f32c(x)
        K = x.shape[-1]
        M = x.numel() // K
        N = w.shape[0]
        y = torch.empty(x.shape[:-1] + (N,), dtype=F32, device=x.device)
        gemm(x, 0, K, 1, w, 0, 1, K, y, 0, N, b, M, N, K)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32c(dy)
        K = x.shape[-1]
        M = x.numel() // K
        N = w.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            gemm(dy, 0, N, 1, w, 0, K, 1, dx, 0, K, None, M, K, N)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            gemm(dy, 0, 1, N, x, 0, K, 1, dw, 0, K, None, N, K, M)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, dtype=F32, device=x.device)
            col_sum(dy, N, M, N, db)
        return dx, dw, db


def linear(x, w, b):
    return LinearFn.apply(x, w, b)


# ------------------------------------------------------------------------------------------------
# GRU chains
# ------------------------------------------------------------------------------------------------
@dataclass
class ChainSpec:
    """One recurrence.  Input-side columns of w_ih: [emb_cols | z_cols] or [x_cols]."""
    emb_cols: Optional[Tuple[int, int]] = None      # (col0, Vin) of w_ih gathered by token id
    ids: Optional[torch.Tensor] = None              # int32 [T,B]
    z_cols: Optional[Tuple[int, int]] = None        # (col0, Zin): time-invariant projection of z_in
    x_cols: Optional[Tuple[int, int]] = None        # (col0, Hin): dense per-step input xin [T,B,Hin]
    h0: Optional[str] = None                        # None (zeros) | "tensor" | "xin0"
    reverse: bool = False
    final: Optional[Tuple[int, int]] = None         # (final buffer index, column offset)
    want_hs: bool = False

    def n_act(self):
        return (self.z_cols is not None) + (self.x_cols is not None) + (self.h0 == "tensor")


class GruGroupFn(torch.autograd.Function):
    """Runs a group of independent GRU chains in one persistent launch.

    apply(specs, B, T, H, final_widths, *tensors); per chain the tensors are
    w_ih, b_ih, w_hh, b_hh, [z_in], [xin], [h0].  Returns (*final_buffers, *hs_of_chains_with_want_hs).
    """

    @staticmethod
    def forward(ctx, specs: List[ChainSpec], B: int, T: int, H: int, final_widths, *tensors):
        dev = tensors[0].device
        require_cuda(*tensors)
        need_grad = any(ctx.needs_input_grad)
        chains = (FnGruChain * len(specs))()
        finals = [torch.empty((B, wd), dtype=F32, device=dev) for wd in final_widths]
        keep = []      # per chain dict of buffers
        pos = 0
        for ci, sp in enumerate(specs):
            w_ih, b_ih, w_hh, b_hh = tensors[pos:pos + 4]
            pos += 4
            z_in = xin = h0 = None
            if sp.z_cols is not None:
                z_in = _f32c(tensors[pos]); pos += 1
            if sp.x_cols is not None:
                xin = _f32c(tensors[pos]); pos += 1
            if sp.h0 == "tensor":
                h0 = _f32c(tensors[pos]); pos += 1
            In = w_ih.shape[1]
            ch = chains[ci]
            d = dict(w_ih=w_ih, b_ih=b_ih, w_hh=w_hh, b_hh=b_hh, z_in=z_in, xin=xin, h0=h0)
            ch.w_hh, ch.b_hh = w_hh.data_ptr(), b_hh.data_ptr()
            if sp.emb_cols is not None:
                c0, Vin = sp.emb_cols
                emb = torch.empty((Vin, 3 * H), dtype=F32, device=dev)
                LIB.call("fn_transpose_f32", _p(w_ih, c0), In, _p(emb), 3 * H, 3 * H, Vin, 0, _st(emb))
                ch.emb, ch.ids = emb.data_ptr(), sp.ids.data_ptr()
                d["emb"] = emb
            if sp.z_cols is not None:
                c0, Zin = sp.z_cols
                proj = torch.empty((B, 3 * H), dtype=F32, device=dev)
                gemm(z_in, 0, Zin, 1, w_ih, c0, 1, In, proj, 0, 3 * H, b_ih, B, 3 * H, Zin)
                ch.proj, ch.proj_ld = proj.data_ptr(), 3 * H
                d["proj"] = proj
            elif sp.x_cols is None:
                ch.proj, ch.proj_ld = b_ih.data_ptr(), 0
            if sp.x_cols is not None:
                c0, Hin = sp.x_cols
                dense = torch.empty((T, B, 3 * H), dtype=F32, device=dev)
                gemm(xin, 0, Hin, 1, w_ih, c0, 1, In, dense, 0, 3 * H, b_ih, T * B, 3 * H, Hin)
                ch.dense = dense.data_ptr()
                d["dense"] = dense
            if sp.h0 == "tensor":
                ch.h0 = h0.data_ptr()
            elif sp.h0 == "xin0":
                ch.h0 = xin.data_ptr()                    # slab 0 of xin [T,B,H]
            ch.reverse = 1 if sp.reverse else 0
            hs = torch.empty((T, B, H), dtype=F32, device=dev)
            ch.hs = hs.data_ptr()
            d["hs"] = hs
            if need_grad:
                gates = torch.empty((T, B, 4 * H), dtype=F32, device=dev)
                ch.gates = gates.data_ptr()
                d["gates"] = gates
            if sp.final is not None:
                fi, fc = sp.final
                ch.h_final = finals[fi].data_ptr() + fc * 4
                ch.h_final_ld = finals[fi].shape[1]
            keep.append(d)
        bar = torch.empty(64 * len(specs), dtype=torch.uint8, device=dev)
        LIB.call("fn_gru_seq_fwd_f32", chains, len(specs), B, T, H, _p(bar), bar.numel(), stream_ptr(dev))
        for d in keep:                                   # forward-only scratch
            d.pop("emb", None); d.pop("proj", None); d.pop("dense", None)
        ctx.specs, ctx.dims, ctx.keep, ctx.n_finals = specs, (B, T, H), keep, len(finals)
        outs = list(finals) + [keep[i]["hs"] for i, sp in enumerate(specs) if sp.want_hs]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        specs, (B, T, H), keep = ctx.specs, ctx.dims, ctx.keep
        if keep is None:
            raise RuntimeError("fadernets_b200: backward through a GRU group a second time -- its saved states were freed "
                               "by the first backward (retain_graph is not supported by this Function)")
        dev = keep[0]["hs"].device
        K3 = 3 * H
        gfinals = [None if g is None else _f32c(g) for g in grads[:ctx.n_finals]]
        ghs_iter = iter(grads[ctx.n_finals:])
        chains = (FnGruChain * len(specs))()
        bufs = []
        for ci, sp in enumerate(specs):
            d, ch = keep[ci], chains[ci]
            ch.w_hh, ch.b_hh = d["w_hh"].data_ptr(), d["b_hh"].data_ptr()
            ch.reverse = 1 if sp.reverse else 0
            ch.hs, ch.gates = d["hs"].data_ptr(), d["gates"].data_ptr()
            if sp.h0 == "tensor":
                ch.h0 = d["h0"].data_ptr()
            elif sp.h0 == "xin0":
                ch.h0 = d["xin"].data_ptr()
            dhs = None
            if sp.want_hs:
                g = next(ghs_iter)
                dhs = None if g is None else _f32c(g)
            if dhs is not None:
                ch.dhs = dhs.data_ptr()
            if sp.final is not None and gfinals[sp.final[0]] is not None:
                gf = gfinals[sp.final[0]]
                ch.dh_final = gf.data_ptr() + sp.final[1] * 4
                ch.dh_final_ld = gf.shape[1]
            b = dict(dgh=torch.empty((T, B, K3), dtype=F32, device=dev),
                     dgin=torch.empty((T, B, H), dtype=F32, device=dev),
                     dh0=torch.empty((B, H), dtype=F32, device=dev),
                     carry=torch.empty((B, H), dtype=F32, device=dev), dhs=dhs)
            ch.dgh, ch.dgin, ch.dh0, ch.dh_carry = (b["dgh"].data_ptr(), b["dgin"].data_ptr(), b["dh0"].data_ptr(),
                                                    b["carry"].data_ptr())
            bufs.append(b)
        bar = torch.empty(64 * len(specs), dtype=torch.uint8, device=dev)
        st = stream_ptr(dev)
        LIB.call("fn_gru_seq_bwd_f32", chains, len(specs), B, T, H, _p(bar), bar.numel(), st)

        out_grads = []
        for ci, sp in enumerate(specs):
            d, b = keep[ci], bufs[ci]
            w_ih, hs, dgh, dgin = d["w_ih"], d["hs"], b["dgh"], b["dgin"]
            In = w_ih.shape[1]
            # ---- recurrent weight: dW_hh = sum_t dgh_t^T h_{t-1}
            dw_hh = torch.empty((K3, H), dtype=F32, device=dev)
            rows = (T - 1) * B
            if not sp.reverse:
                gemm(dgh, B * K3, 1, K3, hs, 0, H, 1, dw_hh, 0, H, None, K3, H, rows)
                h0_rows = 0
            else:
                gemm(dgh, 0, 1, K3, hs, B * H, H, 1, dw_hh, 0, H, None, K3, H, rows)
                h0_rows = (T - 1) * B
            h0t = d["h0"] if sp.h0 == "tensor" else (d["xin"] if sp.h0 == "xin0" else None)
            if h0t is not None:
                gemm(dgh, h0_rows * K3, 1, K3, h0t, 0, H, 1, dw_hh, 0, H, None, K3, H, B, accumulate=True)
            # ---- time sums -> biases and the time-invariant projection
            dproj = torch.empty((B, K3), dtype=F32, device=dev)
            dghsum = torch.empty((B, K3), dtype=F32, device=dev)
            LIB.call("fn_time_sum_f32", _p(dgh), _p(dgin), B, T, H, _p(dproj), _p(dghsum), st)
            db_hh = torch.empty(K3, dtype=F32, device=dev)
            db_ih = torch.empty(K3, dtype=F32, device=dev)
            col_sum(dghsum, K3, B, K3, db_hh)
            col_sum(dproj, K3, B, K3, db_ih)
            # ---- input weight
            covered = sum(c[1] for c in (sp.emb_cols, sp.z_cols, sp.x_cols) if c is not None)
            dw_ih = (torch.empty if covered == In else torch.zeros)((K3, In), dtype=F32, device=dev)
            dz_in = dxin = None
            if sp.emb_cols is not None:
                c0, Vin = sp.emb_cols
                nbytes = LIB.call("fn_emb_grad_scratch_bytes", B, T, H, Vin)
                scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                demb = torch.empty((Vin, K3), dtype=F32, device=dev)
                LIB.call("fn_emb_grad_f32", _p(sp.ids), _p(dgh), _p(dgin), B, T, H, Vin, _p(demb), _p(scratch),
                         nbytes, st)
                LIB.call("fn_transpose_f32", _p(demb), K3, _p(dw_ih, c0), In, Vin, K3, 0, st)
            if sp.z_cols is not None:
                c0, Zin = sp.z_cols
                z_in = d["z_in"]
                gemm(dproj, 0, 1, K3, z_in, 0, Zin, 1, dw_ih, c0, In, None, K3, Zin, B)
                dz_in = torch.empty((B, Zin), dtype=F32, device=dev)
                gemm(dproj, 0, K3, 1, w_ih, c0, In, 1, dz_in, 0, Zin, None, B, Zin, K3)
            if sp.x_cols is not None:
                c0, Hin = sp.x_cols
                xin = d["xin"]
                TB = T * B
                dxin = torch.empty((T, B, Hin), dtype=F32, device=dev)
                gemm(dgh, 0, K3, 1, w_ih, c0, In, 1, dxin, 0, Hin, None, TB, Hin, 2 * H)
                gemm(dgin, 0, H, 1, w_ih, 2 * H * In + c0, In, 1, dxin, 0, Hin, None, TB, Hin, H, accumulate=True)
                gemm(dgh, 0, 1, K3, xin, 0, Hin, 1, dw_ih, c0, In, None, 2 * H, Hin, TB)
                gemm(dgin, 0, 1, H, xin, 0, Hin, 1, dw_ih, 2 * H * In + c0, In, None, H, Hin, TB)
                if sp.h0 == "xin0":
                    LIB.call("fn_add_f32", _p(dxin), _p(b["dh0"]), B * H, st)
            out_grads += [dw_ih, db_ih, dw_hh, db_hh]
            if sp.z_cols is not None:
                out_grads.append(dz_in)
            if sp.x_cols is not None:
                out_grads.append(dxin)
            if sp.h0 == "tensor":
                out_grads.append(b["dh0"])
        ctx.keep = None
        return (None, None, None, None, None) + tuple(out_grads)


# ------------------------------------------------------------------------------------------------
# soft-max heads
# ------------------------------------------------------------------------------------------------
class VocabLogSoftmaxFn(torch.autograd.Function):
    """logits [T,B,V] (time-major) -> log-probs (B,T,V)."""

    @staticmethod
    def forward(ctx, logits_tm):
        logits_tm = _f32c(logits_tm)
        T, B, V = logits_tm.shape
        out = torch.empty((B, T, V), dtype=F32, device=logits_tm.device)
        LIB.call("fn_vocab_logsoftmax_fwd", _p(logits_tm), B, T, V, _p(out), _st(out))
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        B, T, V = out.shape
        dout = _f32c(dout)
        dl = torch.empty((T, B, V), dtype=F32, device=out.device)
        LIB.call("fn_vocab_logsoftmax_bwd", _p(out), _p(dout), B, T, V, _p(dl), _st(dl))
        return dl


class VocabNllFn(torch.autograd.Function):
    """Fused mean NLL of the vocabulary head straight from time-major logits (train fast path)."""

    @staticmethod
    def forward(ctx, logits_tm, target_bm):
        logits_tm = _f32c(logits_tm)
        T, B, V = logits_tm.shape
        dev = logits_tm.device
        target_bm = target_bm.long().contiguous()
        lse = torch.empty((T, B), dtype=F32, device=dev)
        rows = torch.empty((T, B), dtype=F32, device=dev)
        LIB.call("fn_vocab_nll_fwd", _p(logits_tm), _p(target_bm), B, T, V, None, _p(lse), _p(rows), _st(lse))
        loss = torch.empty((), dtype=F32, device=dev)
        scratch, n = _reduce_scratch(dev)
        LIB.call("fn_sum_f32", _p(rows), T * B, 1.0 / (T * B), _p(loss), _p(scratch), n, _st(loss))
        ctx.save_for_backward(logits_tm, lse, target_bm)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits_tm, lse, target = ctx.saved_tensors
        T, B, V = logits_tm.shape
        dloss = _f32c(dloss)
        dl = torch.empty_like(logits_tm)
        LIB.call("fn_vocab_nll_bwd", _p(logits_tm), _p(lse), _p(target), _p(dloss), 1.0 / (T * B), B, T, V, _p(dl),
                 _st(dl))
        return dl, None


class TimeLogSoftmaxFn(torch.autograd.Function):
    """logits [T,B,C] -> (B,T,C) log-softmax over the TIME axis (the reference's dim=1)."""

    @staticmethod
    def forward(ctx, logits_tm):
        logits_tm = _f32c(logits_tm)
        T, B, Cc = logits_tm.shape
        out = torch.empty((B, T, Cc), dtype=F32, device=logits_tm.device)
        LIB.call("fn_time_logsoftmax_fwd", _p(logits_tm), B, T, Cc, _p(out), _st(out))
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        B, T, Cc = out.shape
        dout = _f32c(dout)
        dl = torch.empty((T, B, Cc), dtype=F32, device=out.device)
        LIB.call("fn_time_logsoftmax_bwd", _p(out), _p(dout), B, T, Cc, _p(dl), _st(dl))
        return dl


class NllMeanFn(torch.autograd.Function):
    """F.nll_loss(logp.view(-1,C), target.view(-1), reduction='mean')."""

    @staticmethod
    def forward(ctx, logp, target):
        logp = _f32c(logp)
        Cc = logp.shape[-1]
        rows = logp.numel() // Cc
        target = target.long().contiguous()
        check_index(target, Cc)
        loss = torch.empty((), dtype=F32, device=logp.device)
        scratch, n = _reduce_scratch(logp.device)
        LIB.call("fn_nll_mean_fwd", _p(logp), _p(target), rows, Cc, _p(loss), _p(scratch), n, _st(loss))
        ctx.save_for_backward(target)
        ctx.shape = logp.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (target,) = ctx.saved_tensors
        Cc = ctx.shape[-1]
        rows = target.numel()
        dloss = _f32c(dloss)
        dlogp = torch.empty(ctx.shape, dtype=F32, device=target.device)
        LIB.call("fn_nll_mean_bwd", _p(target), rows, Cc, _p(dloss), _p(dlogp), 1, _st(dlogp))
        return dlogp, None


# ------------------------------------------------------------------------------------------------
# latent block
# ------------------------------------------------------------------------------------------------
def _latent_scratch(B, Z, K, dev):
    """Caller-owned scratch of the two-stage latent reductions (fn_latent_scratch_bytes)."""
    nb = int(LIB.call("fn_latent_scratch_bytes", B, Z, K))
    return torch.empty(max(nb, 16), dtype=torch.uint8, device=dev), nb


class LatentHeadFn(torch.autograd.Function):
    """scale = exp(pre);  z = mu + scale * eps."""

    @staticmethod
    def forward(ctx, mu, pre, eps):
        mu, pre, eps = _f32c(mu), _f32c(pre), _f32c(eps)
        scale, z = torch.empty_like(pre), torch.empty_like(pre)
        LIB.call("fn_reparam_fwd", _p(mu), _p(pre), _p(eps), pre.numel(), _p(scale), _p(z), _st(z))
        ctx.save_for_backward(eps, scale)
        return scale, z

    @staticmethod
    def backward(ctx, dscale, dz):
        eps, scale = ctx.saved_tensors
        dz = None if dz is None else _f32c(dz)
        dscale = None if dscale is None else _f32c(dscale)
        dmu, dpre = torch.empty_like(scale), torch.empty_like(scale)
        LIB.call("fn_reparam_bwd", _p(dz), None, _p(dscale), _p(eps), _p(scale), scale.numel(), _p(dmu), _p(dpre),
                 _st(dmu))
        return dmu, dpre, None


class ExpFn(torch.autograd.Function):
    """scale = exp(pre)  (the `.exp_()` of the variance heads, gmm_model.py:86,91)."""

    @staticmethod
    def forward(ctx, pre):
        pre = _f32c(pre)
        scale = torch.empty_like(pre)
        LIB.call("fn_reparam_fwd", None, _p(pre), None, pre.numel(), _p(scale), None, _st(pre))
        ctx.save_for_backward(scale)
        return scale

    @staticmethod
    def backward(ctx, dscale):
        (scale,) = ctx.saved_tensors
        dscale = _f32c(dscale)
        dpre = torch.empty_like(scale)
        LIB.call("fn_reparam_bwd", None, None, _p(dscale), None, _p(scale), scale.numel(), None, _p(dpre), _st(scale))
        return dpre


class QyXFn(torch.autograd.Function):
    """approx_qy_x: (logLogit, qy_x, y)."""

    @staticmethod
    def forward(ctx, z, mu_lookup, logvar_lookup):
        z, mu_lookup, logvar_lookup = _f32c(z), _f32c(mu_lookup), _f32c(logvar_lookup)
        B, Z = z.shape
        K = mu_lookup.shape[0]
        dev = z.device
        ll = torch.empty((B, K), dtype=F32, device=dev)
        qy = torch.empty((B, K), dtype=F32, device=dev)
        y = torch.empty((B,), dtype=torch.int64, device=dev)
        LIB.call("fn_qy_x_fwd", _p(z), _p(mu_lookup), _p(logvar_lookup), B, Z, K, _p(ll), _p(qy), _p(y), _st(z))
        ctx.save_for_backward(z, mu_lookup, logvar_lookup, qy)
        ctx.mark_non_differentiable(y)
        return ll, qy, y

    @staticmethod
    def backward(ctx, dll, dqy, _dy):
        z, mul, lvl, qy = ctx.saved_tensors
        B, Z = z.shape
        K = mul.shape[0]
        dll = None if dll is None else _f32c(dll)
        dqy = None if dqy is None else _f32c(dqy)
        dz, dmul = torch.empty_like(z), torch.empty_like(mul)
        ws, nb = _latent_scratch(B, Z, K, z.device)
        LIB.call("fn_qy_x_bwd", _p(z), _p(mul), _p(lvl), _p(qy), _p(dll), _p(dqy), B, Z, K, _p(dz), _p(dmul), _p(ws), nb, _st(z))
        return dz, dmul, None


class GmKlFn(torch.autograd.Function):
    """-> tensor[3] = (kld_lat, kld_cls, label_clf); mode 0 unsupervised, 1 supervised."""

    @staticmethod
    def forward(ctx, mu, scale, mu_lookup, logvar_lookup, qy, ll, y_label, mode):
        mu, scale, mu_lookup, logvar_lookup, qy, ll = map(_f32c, (mu, scale, mu_lookup, logvar_lookup, qy, ll))
        B, Z = mu.shape
        K = mu_lookup.shape[0]
        if y_label is not None:
            y_label = y_label.long().contiguous()
            check_index(y_label, K)
        out = torch.empty(3, dtype=F32, device=mu.device)
        ws, nb = _latent_scratch(B, Z, K, mu.device)
        LIB.call("fn_gm_kl_fwd", _p(mu), _p(scale), _p(mu_lookup), _p(logvar_lookup), _p(qy), _p(ll), _p(y_label),
                 mode, B, Z, K, _p(out), _p(ws), nb, _st(out))
        ctx.save_for_backward(mu, scale, mu_lookup, logvar_lookup, qy, ll)
        ctx.y_label, ctx.mode = y_label, mode
        return out

    @staticmethod
    def backward(ctx, dout):
        mu, scale, mul, lvl, qy, ll = ctx.saved_tensors
        B, Z = mu.shape
        K = mul.shape[0]
        dout = _f32c(dout)
        dmu, dsc = torch.empty_like(mu), torch.empty_like(scale)
        dqy, dll, dmul = torch.empty_like(qy), torch.empty_like(ll), torch.empty_like(mul)
        ws, nb = _latent_scratch(B, Z, K, mu.device)
        LIB.call("fn_gm_kl_bwd", _p(mu), _p(scale), _p(mul), _p(lvl), _p(qy), _p(ll), _p(ctx.y_label), ctx.mode,
                 B, Z, K, _p(dout), _p(dmu), _p(dsc), _p(dqy), _p(dll), _p(dmul), _p(ws), nb, _st(mu))
        return dmu, dsc, dmul, None, dqy, dll, None, None


class StdKlFn(torch.autograd.Function):
    """mean_{B,Z} KL(N(mu, scale) || N(0,1))."""

    @staticmethod
    def forward(ctx, mu, scale):
        mu, scale = _f32c(mu), _f32c(scale)
        out = torch.empty((), dtype=F32, device=mu.device)
        ws, nb = _latent_scratch(mu.shape[0], mu.shape[-1], 1, mu.device)
        LIB.call("fn_std_kl_fwd", _p(mu), _p(scale), mu.numel(), _p(out), _p(ws), nb, _st(out))
        ctx.save_for_backward(mu, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        mu, scale = ctx.saved_tensors
        dout = _f32c(dout)
        dmu, dsc = torch.empty_like(mu), torch.empty_like(scale)
        LIB.call("fn_std_kl_bwd", _p(mu), _p(scale), mu.numel(), _p(dout), _p(dmu), _p(dsc), _st(mu))
        return dmu, dsc


class LatentRegFn(torch.autograd.Function):
    """Pati et al. pairwise regulariser on latent dim 0; attr is a float64 device vector (B,)."""

    @staticmethod
    def forward(ctx, z, attr):
        z = _f32c(z)
        B, Z = z.shape
        dev = z.device
        loss = torch.empty((), dtype=F32, device=dev)
        dz0 = torch.empty(B, dtype=F32, device=dev)
        rows = torch.empty(B, dtype=F32, device=dev)
        LIB.call("fn_latent_reg_fwd", _p(z), Z, _p(attr), B, _p(loss), _p(dz0), _p(rows), _st(z))
        ctx.save_for_backward(dz0)
        ctx.shape = (B, Z)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dz0,) = ctx.saved_tensors
        B, Z = ctx.shape
        dloss = _f32c(dloss)
        dz = torch.empty((B, Z), dtype=F32, device=dz0.device)
        LIB.call("fn_latent_reg_bwd", _p(dz0), _p(dloss), B, Z, _p(dz), _st(dz))
        return dz, None


# ------------------------------------------------------------------------------------------------
# small pieces of the sibling models (MusicAttrFaderNets discriminator heads, adversarial MSE)
# ------------------------------------------------------------------------------------------------
class ReluMaskFn(torch.autograd.Function):
    """y = relu(x) * mask -- `dropout(relu(x))` with the keep mask (already divided by 1 - p) drawn by the caller."""

    @staticmethod
    def forward(ctx, x, mask):
        x = _f32c(x)
        mask = None if mask is None else _f32c(mask)
        y = torch.empty_like(x)
        LIB.call("fn_relu_mask_fwd", _p(x), _p(mask), _p(y), x.numel(), _st(x))
        ctx.save_for_backward(x, mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mask = ctx.saved_tensors
        dx = torch.empty_like(x)
        LIB.call("fn_relu_mask_bwd", _p(x), _p(mask), _p(_f32c(dy)), _p(dx), x.numel(), _st(x))
        return dx, None


class GradReverseFn(torch.autograd.Function):
    """ReverseLayerF (model_v2.py:426-435): identity forward, gradient times -alpha."""

    @staticmethod
    def forward(ctx, x, alpha=1.0):
        ctx.alpha = float(alpha)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, dy):
        dy = _f32c(dy)
        dx = torch.empty_like(dy)
        LIB.call("fn_scale_f32", _p(dy), _p(dx), -ctx.alpha, dy.numel(), _st(dy))
        return dx, None


class MseMeanFn(torch.autograd.Function):
    """torch.nn.MSELoss(reduction='mean')(x, y) over a batch vector (trainer_fader.py:107-108); no gradient into y."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.shape = x.shape
        x, y = _f32c(x).reshape(-1), _f32c(y).reshape(-1)
        loss = torch.empty((), dtype=F32, device=x.device)
        LIB.call("fn_mse_mean_fwd", _p(x), _p(y), x.numel(), _p(loss), _st(x))
        ctx.save_for_backward(x, y)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, y = ctx.saved_tensors
        dx = torch.empty_like(x)
        LIB.call("fn_mse_mean_bwd", _p(x), _p(y), x.numel(), _p(_f32c(dloss)), _p(dx), _st(x))
        return dx.view(ctx.shape), None
