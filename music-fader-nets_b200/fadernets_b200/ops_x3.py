"""autograd glue of the bf16x3 mode: the fp32 parity bar (1e-3 relative; BASELINE config 2) ON the tensor cores.

Same graph as fadernets_b200.ops_bf16.  What changes: every fp32 value that feeds a T-scale product is carried as two
bf16 planes, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa bits together), and every product runs as
hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator (fn_gru_seq_*_bf16x3, fn_tc_gemm_bf16x3; include/fadernets_b200.h).
A "split" activation is a bf16 tensor whose last dimension is [hi | lo] (2 x width); everything that is not an operand of
a tensor-core product (gate pre-activations from the input side, incoming state gradients, all GEMM outputs, parameters,
the latent block, reductions) is plain fp32.  torch only allocates; no torch operator does arithmetic here.

Gradients of split activations.  autograd insists that a gradient has the shape and dtype of the tensor it belongs to; the
gradient wrt the VALUE hi + lo of a split tensor [..., 2W] (bf16) is fp32 [..., W] -- the same number of bytes.  It
therefore travels through autograd as the fp32 buffer re-viewed as bf16 [..., 2W] (`as_split_grad` / `from_split_grad`):
producers and consumers of split tensors are only the Functions of this module, and every split tensor has exactly one
consumer in the models (autograd never has to add two such buffers).
"""
from __future__ import annotations

from typing import List

import torch

from ._lib import LIB, FnGruChainBf16, require_cuda, stream_ptr
from .ops import F32, ChainSpec, _f32c, _p, _st, col_sum, gemm
from .ops_bf16 import BF16, GruGroupBf16, _pad32, onehot_bf16, plan_splits, r8


def as_split_grad(g_f32: torch.Tensor) -> torch.Tensor:
    """fp32 gradient [..., W] of a split tensor -> the bf16 [..., 2W] view autograd expects."""
    return g_f32.view(BF16)


def from_split_grad(g: torch.Tensor) -> torch.Tensor:
    """Inverse of as_split_grad (a materialised all-zero gradient is all-zero either way)."""
    g = g if g.is_contiguous() else g.contiguous()
    return g.view(F32) if g.dtype == BF16 else _f32c(g)


def split_bf16(src: torch.Tensor, rows: int, cols: int, s_r: int, s_c: int, off: int = 0, triple: bool = False):
    """(hi, lo) bf16 planes of the fp32 matrix view src[off + r*s_r + c*s_c]: [rows][2*r8(cols)] = [hi | lo], or
    (triple: the K axis of a recurrent weight) [rows][3*r8(cols)] = [hi | hi | lo].  Returns (tensor, ld, lo_off)."""
    cp = r8(cols)
    n = 3 if triple else 2
    dst = (torch.zeros if cp != cols else torch.empty)((rows, n * cp), dtype=BF16, device=src.device)
    lo_off = 2 * cp if triple else cp
    LIB.call("fn_split_bf16", _p(src, off), s_r, s_c, _p(dst), n * cp, rows, cols, lo_off, cp if triple else -1, _st(dst))
    return dst, n * cp, lo_off


def tc_gemm_x3(A, a_off, lda, a_lo, a_mn, B, b_off, ldb, b_lo, b_mn, Cm, c_off, ldc, bias, M, N, K, accumulate=False):
    """C[M][N] fp32 (+)= A * B (+ bias) over hi / lo planes (a_lo / b_lo: element offset of the lo plane, 0 = exact operand)."""
    nprod = 1 + (1 if a_lo else 0) + (1 if b_lo else 0)
    nb, ws = 0, None
    splits = plan_splits(M, N, K, nprod)
    if splits > 1:
        nb = LIB.call("fn_tc_gemm_splitk_ws_bytes", M, N, splits)
        ws = torch.empty(nb, dtype=torch.uint8, device=Cm.device)
    LIB.call("fn_tc_gemm_bf16x3", _p(A, a_off), lda, a_lo, a_mn, _p(B, b_off), ldb, b_lo, b_mn, _p(Cm, c_off), ldc,
             1 if Cm.dtype == BF16 else 0, _p(bias), M, N, K, 1 if accumulate else 0, splits, _p(ws), nb, _st(Cm))


class LinearX3Fn(torch.autograd.Function):
    """y = x W^T + b for a split activation x [..., 2K] (K % 8 == 0) and fp32 master weights W [N,K]; y, dx, dW, db fp32."""

    @staticmethod
    def forward(ctx, x, w, b):
        require_cuda(x, w)
        assert x.dtype == BF16 and x.is_contiguous() and x.shape[-1] % 16 == 0, "bf16x3 linear: split input [..., 2K]"
        K = x.shape[-1] // 2
        M = x.numel() // (2 * K)
        N = w.shape[0]
        ws, ldw, wlo = split_bf16(w, N, K, K, 1)
        y = torch.empty(x.shape[:-1] + (N,), dtype=F32, device=x.device)
        tc_gemm_x3(x, 0, 2 * K, K, 0, ws, 0, ldw, wlo, 0, y, 0, N, b, M, N, K)
        ctx.save_for_backward(x, ws)
        ctx.meta = (M, N, K, ldw, wlo, x.shape, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, ws = ctx.saved_tensors
        M, N, K, ldw, wlo, xshape, has_bias = ctx.meta
        dev = x.device
        dyf = _f32c(dy)
        dys, ldy, ylo = split_bf16(dyf, M, N, N, 1)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(xshape[:-1] + (K,), dtype=F32, device=dev)            # gradient wrt the VALUE hi + lo
            tc_gemm_x3(dys, 0, ldy, ylo, 0, ws, 0, ldw, wlo, 1, dx, 0, K, None, M, K, N)
            dx = as_split_grad(dx)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), dtype=F32, device=dev)
            tc_gemm_x3(dys, 0, ldy, ylo, 1, x, 0, 2 * K, K, 1, dw, 0, K, None, N, K, M)
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, dtype=F32, device=dev)
            col_sum(dyf, N, M, N, db)
        return dx, dw, db


def linear_x3(x, w, b):
    return LinearX3Fn.apply(x, w, b)


def _finish_chain_grads_x3(sp, d, dg, dh0, B, T, H, dxin_done=False):
    """ops_bf16._finish_chain_grads over the split gate-gradient stream dg [T][B][8H] (hi planes | lo planes)."""
    dev = dg.device
    st = stream_ptr(dev)
    K3, H4, H8, TB = 3 * H, 4 * H, 8 * H, T * B
    w_ih, hsx = d["w_ih"], d["hsx"]
    In = w_ih.shape[1]
    dw_hh = torch.empty((K3, H), dtype=F32, device=dev)
    hoff = B * 2 * H if sp.reverse else 0
    tc_gemm_x3(dg, 0, H8, H4, 1, hsx, hoff, 2 * H, H, 1, dw_hh, 0, H, None, 2 * H, H, TB)
    tc_gemm_x3(dg, K3, H8, H4, 1, hsx, hoff, 2 * H, H, 1, dw_hh, 2 * H * H, H, None, H, H, TB)
    dproj = torch.empty((B, K3), dtype=F32, device=dev)
    dghsum = torch.empty((B, K3), dtype=F32, device=dev)
    LIB.call("fn_time_sum_bf16x3", _p(dg), B, T, H, _p(dproj), _p(dghsum), st)
    db_hh = torch.empty(K3, dtype=F32, device=dev)
    db_ih = torch.empty(K3, dtype=F32, device=dev)
    col_sum(dghsum, K3, B, K3, db_hh)
    col_sum(dproj, K3, B, K3, db_ih)
    covered = sum(c[1] for c in (sp.emb_cols, sp.z_cols, sp.x_cols) if c is not None)
    dw_ih = (torch.empty if covered == In else torch.zeros)((K3, In), dtype=F32, device=dev)
    dz_in = dxin = None
    if sp.emb_cols is not None:
        c0, Vin = sp.emb_cols
        oh = onehot_bf16(sp.ids, Vin)                                  # exact in bf16: one plane
        tc_gemm_x3(dg, 0, H8, H4, 1, oh, 0, r8(Vin), 0, 1, dw_ih, c0, In, None, K3, Vin, TB)
    if sp.z_cols is not None:
        c0, Zin = sp.z_cols
        z_in = d["z_in"]
        gemm(dproj, 0, 1, K3, z_in, 0, Zin, 1, dw_ih, c0, In, None, K3, Zin, B)
        dz_in = torch.empty((B, Zin), dtype=F32, device=dev)
        gemm(dproj, 0, K3, 1, w_ih, c0, In, 1, dz_in, 0, Zin, None, B, Zin, K3)
    if sp.x_cols is not None:
        c0, Hin = sp.x_cols
        xin = d["xin"]
        wis, ldw, wlo = d["w_ih_s"]
        if not dxin_done:                                              # (the wavefront computes it per segment)
            dxin = torch.empty((T, B, Hin), dtype=F32, device=dev)
            tc_gemm_x3(dg, 0, H8, H4, 0, wis, c0, ldw, wlo, 1, dxin, 0, Hin, None, TB, Hin, K3)
            if sp.h0 == "xin0":
                LIB.call("fn_add_f32", _p(dxin), _p(dh0), B * Hin, st)     # xin[0] is also this chain's initial state
        tc_gemm_x3(dg, 0, H8, H4, 1, xin, 0, 2 * Hin, Hin, 1, dw_ih, c0, In, None, K3, Hin, TB)
    out = [dw_ih, db_ih, dw_hh, db_hh]
    if sp.z_cols is not None:
        out.append(dz_in)
    if sp.x_cols is not None:
        out.append(None if dxin is None else as_split_grad(dxin))
    if sp.h0 == "tensor":
        out.append(dh0)
    return out


class GruGroupX3Fn(torch.autograd.Function):
    """ops_bf16.GruGroupBf16Fn in bf16x3 mode.  hs outputs are SPLIT bf16 [T,B,2H] views of the chain's [T+1,B,2H] slab
    buffer; a dense input `xin` (and its `h0 == "xin0"`) is expected split, [T,B,2Hin]; hs gradients arrive as fp32 [T,B,H]
    in the bf16 view described in the module docstring."""

    @staticmethod
    def forward(ctx, specs: List[ChainSpec], B: int, T: int, H: int, final_widths, *tensors):
        dev = tensors[0].device
        require_cuda(*tensors)
        need_grad = any(ctx.needs_input_grad)
        n = len(specs)
        chains = (FnGruChainBf16 * n)()
        finals = [torch.empty((B, wd), dtype=F32, device=dev) for wd in final_widths]
        keep, tmp = [], []
        pos = 0
        K3 = 3 * H
        st = stream_ptr(dev)
        for ci, sp in enumerate(specs):
            w_ih, b_ih, w_hh, b_hh = tensors[pos:pos + 4]
            pos += 4
            z_in = xin = h0 = None
            if sp.z_cols is not None:
                z_in = _f32c(tensors[pos]); pos += 1
            if sp.x_cols is not None:
                xin = tensors[pos]; pos += 1
                assert xin.dtype == BF16 and xin.is_contiguous(), "bf16x3 GRU path: dense input must be a contiguous split tensor"
            if sp.h0 == "tensor":
                h0 = _f32c(tensors[pos]); pos += 1
            In = w_ih.shape[1]
            ch = chains[ci]
            d = dict(w_ih=w_ih, w_hh=w_hh, z_in=z_in, xin=xin, spec=sp)
            whs, _, _ = split_bf16(w_hh, K3, H, H, 1, triple=True)          # [3H][hi | hi | lo]
            tmp.append(whs)
            ch.w_hh, ch.b_hh = whs.data_ptr(), b_hh.data_ptr()
            if sp.emb_cols is not None:
                c0, Vin = sp.emb_cols
                emb = torch.empty((Vin, K3), dtype=F32, device=dev)       # fp32 W_ih[:, c0:c0+Vin]^T
                LIB.call("fn_transpose_f32", _p(w_ih, c0), In, _p(emb), K3, K3, Vin, 0, st)
                ch.emb, ch.ids = emb.data_ptr(), sp.ids.data_ptr()
                tmp.append(emb)
            if sp.z_cols is not None:
                c0, Zin = sp.z_cols
                proj = torch.empty((B, K3), dtype=F32, device=dev)
                gemm(z_in, 0, Zin, 1, w_ih, c0, 1, In, proj, 0, K3, b_ih, B, K3, Zin)
                ch.proj, ch.proj_ld = proj.data_ptr(), K3
                tmp.append(proj)
            elif sp.x_cols is None:
                ch.proj, ch.proj_ld = b_ih.data_ptr(), 0
            if sp.x_cols is not None:
                c0, Hin = sp.x_cols
                assert xin.shape[-1] == 2 * Hin
                wis = split_bf16(w_ih, K3, In, In, 1)
                d["w_ih_s"] = wis
                dense = torch.empty((T, B, K3), dtype=F32, device=dev)
                tc_gemm_x3(xin, 0, 2 * Hin, Hin, 0, wis[0], c0, wis[1], wis[2], 0, dense, 0, K3, b_ih, T * B, K3, Hin)
                ch.dense = dense.data_ptr()
                tmp.append(dense)
            hsx = torch.empty((T + 1, B, 2 * H), dtype=BF16, device=dev)
            init = hsx[T if sp.reverse else 0]
            if sp.h0 == "tensor":
                LIB.call("fn_split_bf16", _p(h0), H, 1, _p(init), 2 * H, B, H, H, -1, st)
            elif sp.h0 == "xin0":
                init.copy_(xin[0])
            else:
                init.zero_()
            ch.hsx = hsx.data_ptr()
            ch.reverse = 1 if sp.reverse else 0
            d["hsx"] = hsx
            if need_grad:
                gates = torch.empty((T, _pad32(B), 8 * H), dtype=BF16, device=dev)
                ch.gates = gates.data_ptr()
                d["gates"] = gates
            if sp.final is not None:
                fi, fc = sp.final
                ch.h_final = finals[fi].data_ptr() + fc * 4
                ch.h_final_ld = finals[fi].shape[1]
            keep.append(d)
        bar = torch.empty(64 * n, dtype=torch.uint8, device=dev)
        LIB.call("fn_gru_seq_fwd_bf16x3", chains, n, B, T, H, _p(bar), bar.numel(), st)
        del tmp
        ctx.specs, ctx.dims, ctx.keep, ctx.n_finals = specs, (B, T, H), keep, len(finals)
        outs = list(finals)
        for i, sp in enumerate(specs):
            if sp.want_hs:
                assert not sp.reverse
                outs.append(keep[i]["hsx"][1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        specs, (B, T, H), keep = ctx.specs, ctx.dims, ctx.keep
        if keep is None:
            raise RuntimeError("fadernets_b200: backward through a GRU group a second time -- its saved states were freed "
                               "by the first backward (retain_graph is not supported by this Function)")
        dev = keep[0]["hsx"].device
        K3 = 3 * H
        n = len(specs)
        gfinals = [None if g is None else _f32c(g) for g in grads[:ctx.n_finals]]
        ghs_iter = iter(grads[ctx.n_finals:])
        chains = (FnGruChainBf16 * n)()
        bufs = []
        for ci, sp in enumerate(specs):
            d, ch = keep[ci], chains[ci]
            wht, _, _ = split_bf16(d["w_hh"], H, K3, 1, H, triple=True)      # W_hh^T [H][hi | hi | lo] (9H columns)
            ch.w_hh_t = wht.data_ptr()
            ch.reverse = 1 if sp.reverse else 0
            ch.hsx, ch.gates = d["hsx"].data_ptr(), d["gates"].data_ptr()
            dhs = None
            if sp.want_hs:
                g = next(ghs_iter)
                if g is not None:
                    dhs = from_split_grad(g)                              # fp32 [T,B,H]
                    assert dhs.shape[-1] == H
                    ch.dhs, ch.dhs_f32 = dhs.data_ptr(), 1
            if sp.final is not None and gfinals[sp.final[0]] is not None:
                gf = gfinals[sp.final[0]]
                ch.dh_final = gf.data_ptr() + sp.final[1] * 4
                ch.dh_final_ld = gf.shape[1]
            b = dict(dg=torch.empty((T, B, 8 * H), dtype=BF16, device=dev), dh0=torch.empty((B, H), dtype=F32, device=dev),
                     dhs=dhs, wht=wht)
            ch.dg, ch.dh0 = b["dg"].data_ptr(), b["dh0"].data_ptr()
            bufs.append(b)
        bar = torch.empty(64 * n, dtype=torch.uint8, device=dev)
        LIB.call("fn_gru_seq_bwd_bf16x3", chains, n, B, T, H, _p(bar), bar.numel(), stream_ptr(dev))
        out_grads = []
        for ci, sp in enumerate(specs):
            out_grads += _finish_chain_grads_x3(sp, keep[ci], bufs[ci]["dg"], bufs[ci]["dh0"], B, T, H)
        ctx.keep = None
        return (None, None, None, None, None) + tuple(out_grads)


class GruGroupX3(GruGroupBf16):
    """GruGroupX3Fn for any batch size (chain groups of <= 256 sequences, like ops_bf16.GruGroupBf16)."""
    FN = GruGroupX3Fn


class DecoderStackX3Fn(torch.autograd.Function):
    """ops_bf16.DecoderStackBf16Fn in bf16x3 mode: the four decoder recurrences as a wavefront (cell 2 one time segment behind
    cell 1 in shared launches).  Same arguments; returns SPLIT (hs_r, hs_n, hs_g2) [T,B,2H]."""

    @staticmethod
    def forward(ctx, specs, B: int, T: int, H: int, *tensors):
        from .ops_bf16 import _segments
        dev = tensors[0].device
        require_cuda(*tensors)
        need_grad = any(ctx.needs_input_grad)
        K3, H2, H8 = 3 * H, 2 * H, 8 * H
        st = stream_ptr(dev)
        keep = []
        for ci, sp in enumerate(specs):
            w_ih, b_ih, w_hh, b_hh, z_in, h0 = tensors[6 * ci:6 * ci + 6]
            z_in, h0 = _f32c(z_in), _f32c(h0)
            In = w_ih.shape[1]
            c0, Vin = sp.emb_cols
            zc0, Zin = sp.z_cols
            d = dict(w_ih=w_ih, w_hh=w_hh, b_hh=b_hh, z_in=z_in, xin=None, spec=sp)
            d["w_hh_s"] = split_bf16(w_hh, K3, H, H, 1, triple=True)[0]
            emb = torch.empty((Vin, K3), dtype=F32, device=dev)
            LIB.call("fn_transpose_f32", _p(w_ih, c0), In, _p(emb), K3, K3, Vin, 0, st)
            d["emb"] = emb
            proj = torch.empty((B, K3), dtype=F32, device=dev)
            gemm(z_in, 0, Zin, 1, w_ih, zc0, 1, In, proj, 0, K3, b_ih, B, K3, Zin)
            d["proj"] = proj
            hsx = torch.empty((T + 1, B, H2), dtype=BF16, device=dev)
            LIB.call("fn_split_bf16", _p(h0), H, 1, _p(hsx), H2, B, H, H, -1, st)
            d["hsx"] = hsx
            d["gates"] = torch.empty((T, _pad32(B), H8), dtype=BF16, device=dev) if need_grad else None
            keep.append(d)
        w_ih2, b_ih2, w_hh2, b_hh2 = tensors[18:22]
        g = keep[2]
        sp2 = ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)
        d2 = dict(w_ih=w_ih2, w_hh=w_hh2, b_hh=b_hh2, z_in=None, xin=g["hsx"][1:], spec=sp2)
        d2["w_hh_s"] = split_bf16(w_hh2, K3, H, H, 1, triple=True)[0]
        d2["w_ih_s"] = split_bf16(w_ih2, K3, H, H, 1)
        d2["hsx"] = torch.empty((T + 1, B, H2), dtype=BF16, device=dev)
        d2["gates"] = torch.empty((T, _pad32(B), H8), dtype=BF16, device=dev) if need_grad else None
        dense = torch.empty((T, B, K3), dtype=F32, device=dev)
        keep.append(d2)
        wis, ldw, wlo = d2["w_ih_s"]

        S = _segments(T)
        L = T // S
        bar = torch.empty(64 * 4, dtype=torch.uint8, device=dev)
        for k in range(S + 1):
            chains = (FnGruChainBf16 * 4)()
            n = 0
            if k < S:
                t0 = k * L
                for d in keep[:3]:
                    ch = chains[n]; n += 1
                    ch.w_hh, ch.b_hh = d["w_hh_s"].data_ptr(), d["b_hh"].data_ptr()
                    ch.emb, ch.ids = d["emb"].data_ptr(), d["spec"].ids.data_ptr() + t0 * B * 4
                    ch.proj, ch.proj_ld = d["proj"].data_ptr(), K3
                    ch.hsx = d["hsx"].data_ptr() + t0 * B * H2 * 2
                    if need_grad:
                        ch.gates = d["gates"].data_ptr() + t0 * _pad32(B) * H8 * 2
            if k >= 1:
                t0 = (k - 1) * L
                # cell 2's input projection for this segment (fp32): dense[t] = hs_g[t] W_ih2^T + b_ih2, hs_g[t] = slab t+1
                tc_gemm_x3(g["hsx"], (t0 + 1) * B * H2, H2, H, 0, wis, 0, ldw, wlo, 0, dense, t0 * B * K3, K3, b_ih2, L * B, K3, H)
                if k == 1:
                    d2["hsx"][0].copy_(g["hsx"][1])                 # hx[1] <- the new hx[0] at step 0 (gmm_model.py:134-135)
                ch = chains[n]; n += 1
                ch.w_hh, ch.b_hh = d2["w_hh_s"].data_ptr(), b_hh2.data_ptr()
                ch.dense = dense.data_ptr() + t0 * B * K3 * 4
                ch.hsx = d2["hsx"].data_ptr() + t0 * B * H2 * 2
                if need_grad:
                    ch.gates = d2["gates"].data_ptr() + t0 * _pad32(B) * H8 * 2
            LIB.call("fn_gru_seq_fwd_bf16x3", chains, n, B, L, H, _p(bar), bar.numel(), st)
        for d in keep:
            d.pop("emb", None); d.pop("proj", None); d.pop("w_hh_s", None)
        ctx.keep, ctx.dims, ctx.S = keep, (B, T, H), S
        return keep[0]["hsx"][1:], keep[1]["hsx"][1:], d2["hsx"][1:]

    @staticmethod
    def backward(ctx, g_r, g_n, g_2):
        keep, (B, T, H), S = ctx.keep, ctx.dims, ctx.S
        if keep is None:
            raise RuntimeError("fadernets_b200: backward through the decoder stack a second time -- its saved states were "
                               "freed by the first backward (retain_graph is not supported by this Function)")
        dev = keep[0]["hsx"].device
        K3, H2, H4, H8 = 3 * H, 2 * H, 4 * H, 8 * H
        L = T // S
        st = stream_ptr(dev)
        dhs = [None if gr is None else from_split_grad(gr) for gr in (g_r, g_n)]
        dhs += [torch.empty((T, B, H), dtype=F32, device=dev), None if g_2 is None else from_split_grad(g_2)]
        whts = [split_bf16(d["w_hh"], H, K3, 1, H, triple=True)[0] for d in keep]     # W_hh^T [H][hi | hi | lo]
        dgs = [torch.empty((T, B, H8), dtype=BF16, device=dev) for _ in keep]
        dh0 = [[torch.empty((B, H), dtype=F32, device=dev) for _ in range(2)] for _ in keep]
        last = [None] * 4
        bar = torch.empty(64 * 4, dtype=torch.uint8, device=dev)
        d2 = keep[3]
        wis, ldw, wlo = d2["w_ih_s"]

        def fill(ch, ci, j, flip):
            d = keep[ci]
            t0 = j * L
            ch.w_hh_t = whts[ci].data_ptr()
            ch.hsx = d["hsx"].data_ptr() + t0 * B * H2 * 2
            ch.gates = d["gates"].data_ptr() + t0 * _pad32(B) * H8 * 2
            ch.dg = dgs[ci].data_ptr() + t0 * B * H8 * 2
            if dhs[ci] is not None:
                ch.dhs = dhs[ci].data_ptr() + t0 * B * H * 4
                ch.dhs_f32 = 1
            if last[ci] is not None:
                ch.dh_final, ch.dh_final_ld = last[ci].data_ptr(), H
            out = dh0[ci][flip]
            ch.dh0 = out.data_ptr()
            return out

        for k in range(S + 1):
            chains = (FnGruChainBf16 * 4)()
            n = 0
            new_last = list(last)
            if k < S:
                new_last[3] = fill(chains[n], 3, S - 1 - k, k & 1); n += 1
            if k >= 1:
                for ci in range(3):
                    new_last[ci] = fill(chains[n], ci, S - k, k & 1); n += 1
            LIB.call("fn_gru_seq_bwd_bf16x3", chains, n, B, L, H, _p(bar), bar.numel(), st)
            last = new_last
            if k < S:
                # gradient wrt cell 1's states of this segment = cell 2's input gradient: dgi W_ih2 (fp32)
                j = S - 1 - k
                t0 = j * L
                tc_gemm_x3(dgs[3], t0 * B * H8, H8, H4, 0, wis, 0, ldw, wlo, 1, dhs[2], t0 * B * H, H, None, L * B, H, K3)
                if j == 0:
                    LIB.call("fn_add_f32", _p(dhs[2]), _p(last[3]), B * H, st)

        out_grads = []
        for ci in range(3):
            out_grads += _finish_chain_grads_x3(keep[ci]["spec"], keep[ci], dgs[ci], last[ci], B, T, H)
        out_grads += _finish_chain_grads_x3(d2["spec"], d2, dgs[3], last[3], B, T, H, dxin_done=True)[:4]
        ctx.keep = None
        return (None, None, None, None) + tuple(out_grads)
