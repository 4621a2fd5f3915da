"""autograd glue of the bf16 tensor-core path (tcgen05): the persistent GRU gate block
(fn_gru_seq_*_bf16) and every batched product (fn_tc_gemm_bf16).

Same graph as fadernets_b200.ops (the fp32 exact-parity path); what changes is the storage of the
T-scale activations -- hidden states, saved gates and gate gradients are bf16, time-major, with the
initial state as an extra slab -- and that every product with a T*B-row operand runs on the tensor
cores with fp32 accumulation.  Parameters, their gradients, the latent block and all reductions stay
fp32.  torch only allocates; no torch operator does arithmetic here.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from ._lib import LIB, FnGruChainBf16, require_cuda, stream_ptr
from .ops import F32, ChainSpec, _f32c, _p, _st, col_sum, gemm

BF16 = torch.bfloat16


def _pad32(n: int) -> int:
    """Rows of a saved-gates slab: the kernels store them in 32-row x 16-column blocks (include/fadernets_b200.h)."""
    return (n + 31) // 32 * 32


def r8(n: int) -> int:
    return (n + 7) // 8 * 8


def cast_bf16(src: torch.Tensor, rows: int, cols: int, s_r: int, s_c: int, off: int = 0, ld_dst: int | None = None):
    """bf16 copy of the fp32 matrix view src[off + r*s_r + c*s_c] -> [rows][ld_dst]."""
    ld = cols if ld_dst is None else ld_dst
    dst = torch.empty((rows, ld), dtype=BF16, device=src.device)
    LIB.call("fn_cast_bf16", _p(src, off), s_r, s_c, _p(dst), ld, rows, cols, _st(dst))
    return dst


def to_bf16_rows(x: torch.Tensor, cols: int):
    """x: [..., cols] fp32 or bf16 -> (bf16 2-D [rows][ld], ld) with ld % 8 == 0 (TMA row pitch)."""
    rows = x.numel() // cols
    if x.dtype == BF16 and cols % 8 == 0 and x.is_contiguous():
        return x.view(rows, cols), cols
    xf = _f32c(x)
    ld = r8(cols)
    return cast_bf16(xf, rows, cols, cols, 1, ld_dst=ld), ld


_SM_SLOTS = 2 * 148          # CTA slots of the 128x128-tile kernel on a B200 (2 per SM)
_SM_PAIRS = 148 // 2         # CTA pairs of the 256x256-tile kernel (one persistent pair per TPC)


def pair_kernel_takes(M: int, N: int, K: int) -> bool:
    """Mirror of fn_tc_gemm2_eligible (csrc/fn_tc_gemm2.cu): shapes that run on the CTA-pair kernel."""
    import os
    if os.environ.get("FN_GEMM_PAIR", "1") == "0":
        return False
    bn = 176 if 256 < N <= 352 else 256           # column-tile width (176: the 342-wide vocabulary)
    return M >= 256 and N > 256 and ((N + bn - 1) // bn * bn - N) * 4 <= N and K >= 64


def plan_splits(M: int, N: int, K: int, nprod: int = 1) -> int:
    """Split-K factor of a tensor-core product with K (x nprod plane products in bf16x3 mode) long and few output tiles (the T*B-row
    weight gradients).  Pair kernel: work items = splits x 256^2 tiles are dealt round-robin to 74 persistent pairs, so the
    factor minimises waves x K-loop time + partial-tile traffic; 128^2 kernel: fill the 296 CTA slots once."""
    pair = pair_kernel_takes(M, N, K)
    K = K * nprod
    if K < 8192:
        return 1
    if pair:
        bn = 176 if 256 < N <= 352 else 256
        tiles = ((M + 255) // 256) * ((N + bn - 1) // bn)
        if tiles >= 2 * _SM_PAIRS:
            return 1
        # cost model (microseconds): waves x (one tile's K loop / s) at ~21 TFLOP/s per pair + 2 us of pipeline fill per wave,
        # + s partial tiles written and read back at ~5 TB/s
        t_tile = 2.0 * 256 * bn * K / 21.0e6
        t_part = M * N * 8 / 5.0e6
        best, best_t = 1, None
        for s_ in range(1, min(32, K // 1024) + 1):
            waves = -(-tiles * s_ // _SM_PAIRS)
            t = waves * (t_tile / s_ + 2.0) + (s_ * t_part if s_ > 1 else 0.0)
            if best_t is None or t < best_t:
                best, best_t = s_, t
        return best
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if tiles >= _SM_SLOTS:
        return 1
    return max(1, min(32, _SM_SLOTS // tiles, K // 2048))


def tc_gemm(A, a_off, lda, a_mn, B, b_off, ldb, b_mn, Cm, c_off, ldc, bias, M, N, K, accumulate=False):
    """C[M][N] (fp32 or bf16 by Cm.dtype) (+)= A * B (+ bias); operand layouts as in fn_tc_gemm_bf16.
    Products with few output tiles and a long K (the T*B-row weight gradients) are split along K."""
    splits = plan_splits(M, N, K)
    if splits > 1:
        nb = LIB.call("fn_tc_gemm_splitk_ws_bytes", M, N, splits)
        ws = torch.empty(nb, dtype=torch.uint8, device=Cm.device)
        LIB.call("fn_tc_gemm_bf16_splitk", _p(A, a_off), lda, a_mn, _p(B, b_off), ldb, b_mn, _p(Cm, c_off), ldc,
                 1 if Cm.dtype == BF16 else 0, _p(bias), M, N, K, 1 if accumulate else 0, splits, _p(ws), nb, _st(Cm))
        return
    LIB.call("fn_tc_gemm_bf16", _p(A, a_off), lda, a_mn, _p(B, b_off), ldb, b_mn, _p(Cm, c_off), ldc,
             1 if Cm.dtype == BF16 else 0, _p(bias), M, N, K, 1 if accumulate else 0, _st(Cm))


def col_sum_bf16(x, ld, rows, cols, out, accumulate=False):
    scratch = torch.empty(64 * cols * 4, dtype=torch.uint8, device=x.device)
    LIB.call("fn_col_sum_bf16", _p(x), ld, rows, cols, _p(out), 1 if accumulate else 0, _p(scratch), scratch.numel(),
             _st(x))


def onehot_bf16(ids_tm: torch.Tensor, V: int):
    """int32 ids [rows...] -> bf16 one-hot [rows][r8(V)]."""
    rows = ids_tm.numel()
    ld = r8(V)
    out = torch.empty((rows, ld), dtype=BF16, device=ids_tm.device)
    LIB.call("fn_ids_to_onehot_bf16", _p(ids_tm), rows, V, ld, _p(out), _st(out))
    return out


# ------------------------------------------------------------------------------------------------
# Linear on the tensor cores
# ------------------------------------------------------------------------------------------------
class LinearBf16Fn(torch.autograd.Function):
    """y = x W^T + b with x [..., K] (bf16 or fp32; K % 8 == 0), W [N,K] fp32 master weights.
    y is fp32 unless out_bf16.  dx is returned in x's dtype; dW, db fp32."""

    @staticmethod
    def forward(ctx, x, w, b, out_bf16=False):
        require_cuda(x, w)
        K = x.shape[-1]
        M = x.numel() // K
        N = w.shape[0]
        xb, ldx = to_bf16_rows(x, K)
        wb = cast_bf16(w, N, K, K, 1, ld_dst=r8(K))
        y = torch.empty(x.shape[:-1] + (N,), dtype=BF16 if out_bf16 else F32, device=x.device)
        tc_gemm(xb, 0, ldx, 0, wb, 0, r8(K), 0, y, 0, N, b, M, N, K)
        ctx.save_for_backward(xb, wb)
        ctx.meta = (M, N, K, ldx, x.dtype, x.shape, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        M, N, K, ldx, xdtype, xshape, has_bias = ctx.meta
        dev = xb.device
        dyb, ldy = to_bf16_rows(dy, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(xshape, dtype=xdtype, device=dev)
            tc_gemm(dyb, 0, ldy, 0, wb, 0, r8(K), 1, dx, 0, K, None, M, K, N)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), dtype=F32, device=dev)
            tc_gemm(dyb, 0, ldy, 1, xb, 0, ldx, 1, dw, 0, K, None, N, K, M)
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, dtype=F32, device=dev)
            col_sum_bf16(dyb, ldy, M, N, db)
        return dx, dw, db, None


def linear_bf16(x, w, b, out_bf16=False):
    return LinearBf16Fn.apply(x, w, b, out_bf16)


def _finish_chain_grads(sp, d, dg, dh0, B, T, H, dxin_done=None):
    """Parameter / input gradients of one chain from its gate-gradient stream dg [T][B][4H] = (dr, dz, dn, dn*r):
    returns [dW_ih, db_ih, dW_hh, db_hh, (dz_in), (dxin), (dh0)] in the order of the chain's inputs."""
    dev = dg.device
    st = stream_ptr(dev)
    K3, H4, TB = 3 * H, 4 * H, T * B
    w_ih, hsx = d["w_ih"], d["hsx"]
    In = w_ih.shape[1]
    # ---- recurrent weight: dW_hh = sum_tau dgh_tau^T (state before that step); rows tau*B+b of dg pair
    # with slab tau (forward chain) / tau+1 (reverse chain) of hsx.  dgh = dg[:, :2H] | dg[:, 3H:].
    dw_hh = torch.empty((K3, H), dtype=F32, device=dev)
    hoff = B * H if sp.reverse else 0
    tc_gemm(dg, 0, H4, 1, hsx, hoff, H, 1, dw_hh, 0, H, None, 2 * H, H, TB)
    tc_gemm(dg, K3, H4, 1, hsx, hoff, H, 1, dw_hh, 2 * H * H, H, None, H, H, TB)
    # ---- time sums -> biases and the time-invariant projection
    dproj = torch.empty((B, K3), dtype=F32, device=dev)
    dghsum = torch.empty((B, K3), dtype=F32, device=dev)
    LIB.call("fn_time_sum_bf16", _p(dg), B, T, H, _p(dproj), _p(dghsum), st)
    db_hh = torch.empty(K3, dtype=F32, device=dev)
    db_ih = torch.empty(K3, dtype=F32, device=dev)
    col_sum(dghsum, K3, B, K3, db_hh)
    col_sum(dproj, K3, B, K3, db_ih)
    # ---- input weight
    covered = sum(c[1] for c in (sp.emb_cols, sp.z_cols, sp.x_cols) if c is not None)
    dw_ih = (torch.empty if covered == In else torch.zeros)((K3, In), dtype=F32, device=dev)
    dz_in = dxin = None
    if sp.emb_cols is not None:
        # autograd of `onehot @ W_ih[:, :Vin]^T`: dW_ih[:, :Vin] = dgi^T onehot, on the tensor cores
        # (the one-hot operand is exact in bf16)
        c0, Vin = sp.emb_cols
        oh = onehot_bf16(sp.ids, Vin)
        tc_gemm(dg, 0, H4, 1, oh, 0, r8(Vin), 1, dw_ih, c0, In, None, K3, Vin, TB)
    if sp.z_cols is not None:
        c0, Zin = sp.z_cols
        z_in = d["z_in"]
        gemm(dproj, 0, 1, K3, z_in, 0, Zin, 1, dw_ih, c0, In, None, K3, Zin, B)
        dz_in = torch.empty((B, Zin), dtype=F32, device=dev)
        gemm(dproj, 0, K3, 1, w_ih, c0, In, 1, dz_in, 0, Zin, None, B, Zin, K3)
    if sp.x_cols is not None:
        c0, Hin = sp.x_cols
        xin, wib = d["xin"], d["w_ih_b"]
        if dxin_done is None:
            dxin = torch.empty((T, B, Hin), dtype=BF16, device=dev)
            tc_gemm(dg, 0, H4, 0, wib, c0, r8(In), 1, dxin, 0, Hin, None, TB, Hin, K3)
            if sp.h0 == "xin0":
                # grad wrt xin[0] also receives the initial-state gradient of this chain
                LIB.call("fn_add_f32_to_bf16", _p(dxin), _p(dh0), B * Hin, st)
        else:
            dxin = dxin_done
        tc_gemm(dg, 0, H4, 1, xin, 0, Hin, 1, dw_ih, c0, In, None, K3, Hin, TB)
    out = [dw_ih, db_ih, dw_hh, db_hh]
    if sp.z_cols is not None:
        out.append(dz_in)
    if sp.x_cols is not None:
        out.append(dxin)
    if sp.h0 == "tensor":
        out.append(dh0)
    return out


# ------------------------------------------------------------------------------------------------
# GRU chains on the tensor cores
# ------------------------------------------------------------------------------------------------
class GruGroupBf16Fn(torch.autograd.Function):
    """Same contract as ops.GruGroupFn (specs, B, T, H, final_widths, *tensors) with bf16 state storage.
    hs outputs are bf16 [T,B,H] views of the chain's [T+1,B,H] slab buffer; a dense input `xin` (and its
    `h0 == "xin0"`) is expected in bf16 [T,B,Hin]."""

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
        for ci, sp in enumerate(specs):
            w_ih, b_ih, w_hh, b_hh = tensors[pos:pos + 4]
            pos += 4
            z_in = xin = h0 = None
            if sp.z_cols is not None:
                z_in = _f32c(tensors[pos]); pos += 1
            if sp.x_cols is not None:
                xin = tensors[pos]; pos += 1
                assert xin.dtype == BF16 and xin.is_contiguous(), "bf16 GRU path: dense input must be contiguous bf16"
            if sp.h0 == "tensor":
                h0 = _f32c(tensors[pos]); pos += 1
            In = w_ih.shape[1]
            ch = chains[ci]
            d = dict(w_ih=w_ih, w_hh=w_hh, z_in=z_in, xin=xin, spec=sp)
            whb = cast_bf16(w_hh, K3, H, H, 1)
            d["w_hh_b"] = whb
            ch.w_hh, ch.b_hh = whb.data_ptr(), b_hh.data_ptr()
            if sp.emb_cols is not None:
                c0, Vin = sp.emb_cols
                emb = cast_bf16(w_ih, Vin, K3, 1, In, off=c0)          # bf16 W_ih[:, c0:c0+Vin]^T : [Vin][3H]
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
                wib = cast_bf16(w_ih, K3, In, In, 1, ld_dst=r8(In))
                d["w_ih_b"] = wib
                dense = torch.empty((T, B, K3), dtype=BF16, device=dev)
                tc_gemm(xin, 0, Hin, 0, wib, c0, r8(In), 0, dense, 0, K3, b_ih, T * B, K3, Hin)
                ch.dense = dense.data_ptr()
                tmp.append(dense)
            hsx = torch.empty((T + 1, B, H), dtype=BF16, device=dev)
            init = hsx[T if sp.reverse else 0]
            if sp.h0 == "tensor":
                LIB.call("fn_cast_bf16", _p(h0), H, 1, _p(init), H, B, H, _st(init))
            elif sp.h0 == "xin0":
                init.copy_(xin[0])
            else:
                init.zero_()
            ch.hsx = hsx.data_ptr()
            ch.reverse = 1 if sp.reverse else 0
            d["hsx"] = hsx
            if need_grad:
                gates = torch.empty((T, _pad32(B), 4 * H), dtype=BF16, device=dev)
                ch.gates = gates.data_ptr()
                d["gates"] = gates
            if sp.final is not None:
                fi, fc = sp.final
                ch.h_final = finals[fi].data_ptr() + fc * 4
                ch.h_final_ld = finals[fi].shape[1]
            keep.append(d)
        bar = torch.empty(64 * n, dtype=torch.uint8, device=dev)
        LIB.call("fn_gru_seq_fwd_bf16", chains, n, B, T, H, _p(bar), bar.numel(), stream_ptr(dev))
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
        K3, H4, TB = 3 * H, 4 * H, T * B
        n = len(specs)
        gfinals = [None if g is None else _f32c(g) for g in grads[:ctx.n_finals]]
        ghs_iter = iter(grads[ctx.n_finals:])
        chains = (FnGruChainBf16 * n)()
        bufs = []
        for ci, sp in enumerate(specs):
            d, ch = keep[ci], chains[ci]
            wht = cast_bf16(d["w_hh"], H, K3, 1, H)                    # W_hh^T [H][3H]
            ch.w_hh_t = wht.data_ptr()
            ch.reverse = 1 if sp.reverse else 0
            ch.hsx, ch.gates = d["hsx"].data_ptr(), d["gates"].data_ptr()
            dhs = None
            if sp.want_hs:
                g = next(ghs_iter)
                if g is not None:
                    dhs = g if g.is_contiguous() else g.contiguous()
                    ch.dhs = dhs.data_ptr()
                    ch.dhs_f32 = 0 if dhs.dtype == BF16 else 1
                    if dhs.dtype not in (BF16, F32):
                        dhs = dhs.float(); ch.dhs = dhs.data_ptr(); ch.dhs_f32 = 1
            if sp.final is not None and gfinals[sp.final[0]] is not None:
                gf = gfinals[sp.final[0]]
                ch.dh_final = gf.data_ptr() + sp.final[1] * 4
                ch.dh_final_ld = gf.shape[1]
            b = dict(dg=torch.empty((T, B, H4), dtype=BF16, device=dev), dh0=torch.empty((B, H), dtype=F32, device=dev),
                     dhs=dhs, wht=wht)
            ch.dg, ch.dh0 = b["dg"].data_ptr(), b["dh0"].data_ptr()
            bufs.append(b)
        bar = torch.empty(64 * n, dtype=torch.uint8, device=dev)
        st = stream_ptr(dev)
        LIB.call("fn_gru_seq_bwd_bf16", chains, n, B, T, H, _p(bar), bar.numel(), st)

        out_grads = []
        for ci, sp in enumerate(specs):
            out_grads += _finish_chain_grads(sp, keep[ci], bufs[ci]["dg"], bufs[ci]["dh0"], B, T, H)
        ctx.keep = None
        return (None, None, None, None, None) + tuple(out_grads)


class GruGroupBf16:
    """GruGroupBf16Fn for any batch size: the tensor-core gate block handles at most 256 rows (two 128-row MMA tiles)
    per chain, so a larger per-GPU batch is cut into independent groups of <= 256 sequences (a recurrence never mixes
    sequences), each run through GruGroupBf16Fn, and the per-sequence outputs are concatenated.  Weight gradients add up
    through autograd.  (The reference's nn.GRU has no batch limit: trainer_gmm.py uses batch_size 128 by default.)"""
    MAX_ROWS = 256
    FN = GruGroupBf16Fn

    @classmethod
    def apply(cls, specs, B, T, H, final_widths, *tensors):
        import dataclasses
        if B <= GruGroupBf16.MAX_ROWS:
            return cls.FN.apply(specs, B, T, H, final_widths, *tensors)
        outs = []
        for lo in range(0, B, GruGroupBf16.MAX_ROWS):
            hi = min(B, lo + GruGroupBf16.MAX_ROWS)
            sub_specs, sub_tensors, pos = [], [], 0
            for sp in specs:
                sub_specs.append(dataclasses.replace(sp, ids=None if sp.ids is None else sp.ids[:, lo:hi].contiguous()))
                sub_tensors += list(tensors[pos:pos + 4]); pos += 4
                if sp.z_cols is not None:
                    sub_tensors.append(tensors[pos][lo:hi]); pos += 1
                if sp.x_cols is not None:
                    sub_tensors.append(tensors[pos][:, lo:hi].contiguous()); pos += 1
                if sp.h0 == "tensor":
                    sub_tensors.append(tensors[pos][lo:hi]); pos += 1
            outs.append(cls.FN.apply(sub_specs, hi - lo, T, H, final_widths, *sub_tensors))
        nf = len(final_widths)
        return tuple(torch.cat([o[i] for o in outs], 0 if i < nf else 1) for i in range(len(outs[0])))


# ------------------------------------------------------------------------------------------------
# decoder stack: sub-decoder r, sub-decoder n, global cell 1 and global cell 2 as one wavefront
# ------------------------------------------------------------------------------------------------
def _segments(T: int):
    """Time segments of the wavefront: cell 2 runs one segment behind cell 1 in the same launch.  S segments cost
    T (S + 1) / S chain steps instead of 2 T; more segments = fewer idle steps in the head / tail launches but more
    launches (FN_WAVEFRONT_SEGMENTS, default 16 where T allows: config 3 measured 43.0 / 42.2 / 42.0 ms per step at 4 / 8 / 16)."""
    import os
    want = int(os.environ.get("FN_WAVEFRONT_SEGMENTS", "16"))
    for S in (16, 8):
        if want >= S and T % S == 0 and T // S >= 32:
            return S
    if T % 4 == 0 and T >= 64:
        return 4
    if T % 2 == 0 and T >= 16:
        return 2
    return 1


class DecoderStackBf16Fn(torch.autograd.Function):
    """The four decoder recurrences of one forward pass (gmm_model.py:100-149): sub-decoder r, sub-decoder n, global
    cell 1 (teacher-forced token gather + z projection) and global cell 2 (dense input = cell 1's states, initial
    state = cell 1's first state).  Cell 2 only depends on cell 1's *earlier* states, so time is cut into S segments
    and launch k runs {r, n, cell 1} on segment k together with {cell 2} on segment k-1 (its input projection for
    that segment is one batched tensor-core GEMM in between): S+1 launches of T/S steps instead of 2 of T, and no
    launch that fills only a quarter of the machine.  Backward mirrors it.

    apply(specs, B, T, H, *tensors): specs = (r, n, g) ChainSpecs (emb + z projection + h0 tensor); tensors = per chain
    w_ih, b_ih, w_hh, b_hh, z_in, h0 for r, n, g, then w_ih, b_ih, w_hh, b_hh of cell 2.
    Returns (hs_r, hs_n, hs_g2): bf16 [T,B,H]."""

    @staticmethod
    def forward(ctx, specs, B: int, T: int, H: int, *tensors):
        dev = tensors[0].device
        require_cuda(*tensors)
        need_grad = any(ctx.needs_input_grad)
        K3, H4 = 3 * H, 4 * H
        st = stream_ptr(dev)
        keep = []
        for ci, sp in enumerate(specs):
            w_ih, b_ih, w_hh, b_hh, z_in, h0 = tensors[6 * ci:6 * ci + 6]
            z_in, h0 = _f32c(z_in), _f32c(h0)
            In = w_ih.shape[1]
            c0, Vin = sp.emb_cols
            zc0, Zin = sp.z_cols
            d = dict(w_ih=w_ih, w_hh=w_hh, b_hh=b_hh, z_in=z_in, xin=None, spec=sp)
            d["w_hh_b"] = cast_bf16(w_hh, K3, H, H, 1)
            d["emb"] = cast_bf16(w_ih, Vin, K3, 1, In, off=c0)
            proj = torch.empty((B, K3), dtype=F32, device=dev)
            gemm(z_in, 0, Zin, 1, w_ih, zc0, 1, In, proj, 0, K3, b_ih, B, K3, Zin)
            d["proj"] = proj
            hsx = torch.empty((T + 1, B, H), dtype=BF16, device=dev)
            LIB.call("fn_cast_bf16", _p(h0), H, 1, _p(hsx), H, B, H, st)
            d["hsx"] = hsx
            d["gates"] = torch.empty((T, _pad32(B), H4), dtype=BF16, device=dev) if need_grad else None
            keep.append(d)
        w_ih2, b_ih2, w_hh2, b_hh2 = tensors[18:22]
        g = keep[2]
        sp2 = ChainSpec(x_cols=(0, H), h0="xin0", want_hs=True)
        d2 = dict(w_ih=w_ih2, w_hh=w_hh2, b_hh=b_hh2, z_in=None, xin=g["hsx"][1:], spec=sp2)
        d2["w_hh_b"] = cast_bf16(w_hh2, K3, H, H, 1)
        d2["w_ih_b"] = cast_bf16(w_ih2, K3, H, H, 1)
        d2["hsx"] = torch.empty((T + 1, B, H), dtype=BF16, device=dev)
        d2["gates"] = torch.empty((T, _pad32(B), H4), dtype=BF16, device=dev) if need_grad else None
        dense = torch.empty((T, B, K3), dtype=BF16, device=dev)
        keep.append(d2)

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
                    ch.w_hh, ch.b_hh = d["w_hh_b"].data_ptr(), d["b_hh"].data_ptr()
                    ch.emb, ch.ids = d["emb"].data_ptr(), d["spec"].ids.data_ptr() + t0 * B * 4
                    ch.proj, ch.proj_ld = d["proj"].data_ptr(), K3
                    ch.hsx = d["hsx"].data_ptr() + t0 * B * H * 2
                    if need_grad:
                        ch.gates = d["gates"].data_ptr() + t0 * _pad32(B) * H4 * 2
            if k >= 1:
                t0 = (k - 1) * L
                # cell 2's input projection for this segment: dense[t] = hs_g[t] W_ih2^T + b_ih2, hs_g[t] = slab t+1
                tc_gemm(g["hsx"], (t0 + 1) * B * H, H, 0, d2["w_ih_b"], 0, H, 0, dense, t0 * B * K3, K3, b_ih2, L * B, K3, H)
                if k == 1:
                    d2["hsx"][0].copy_(g["hsx"][1])                 # hx[1] <- the new hx[0] at step 0 (gmm_model.py:134-135)
                ch = chains[n]; n += 1
                ch.w_hh, ch.b_hh = d2["w_hh_b"].data_ptr(), b_hh2.data_ptr()
                ch.dense = dense.data_ptr() + t0 * B * K3 * 2
                ch.hsx = d2["hsx"].data_ptr() + t0 * B * H * 2
                if need_grad:
                    ch.gates = d2["gates"].data_ptr() + t0 * _pad32(B) * H4 * 2
            LIB.call("fn_gru_seq_fwd_bf16", chains, n, B, L, H, _p(bar), bar.numel(), st)
        for d in keep:
            d.pop("emb", None); d.pop("proj", None)
        ctx.keep, ctx.dims, ctx.S = keep, (B, T, H), S
        return keep[0]["hsx"][1:], keep[1]["hsx"][1:], d2["hsx"][1:]

    @staticmethod
    def backward(ctx, g_r, g_n, g_2):
        keep, (B, T, H), S = ctx.keep, ctx.dims, ctx.S
        if keep is None:
            raise RuntimeError("fadernets_b200: backward through the decoder stack a second time -- its saved states were "
                               "freed by the first backward (retain_graph is not supported by this Function)")
        dev = keep[0]["hsx"].device
        K3, H4 = 3 * H, 4 * H
        L = T // S
        st = stream_ptr(dev)

        def as_dhs(gr):
            if gr is None:
                return None
            gr = gr if gr.is_contiguous() else gr.contiguous()
            return gr if gr.dtype in (BF16, F32) else gr.float()

        dhs = [as_dhs(g_r), as_dhs(g_n), torch.empty((T, B, H), dtype=BF16, device=dev), as_dhs(g_2)]
        whts = [cast_bf16(d["w_hh"], H, K3, 1, H) for d in keep]                     # W_hh^T [H][3H]
        dgs = [torch.empty((T, B, H4), dtype=BF16, device=dev) for _ in keep]
        dh0 = [[torch.empty((B, H), dtype=F32, device=dev) for _ in range(2)] for _ in keep]
        last = [None] * 4                                                             # dh0 of the later segment
        bar = torch.empty(64 * 4, dtype=torch.uint8, device=dev)
        d2 = keep[3]

        def fill(ch, ci, j, flip):
            d = keep[ci]
            t0 = j * L
            ch.w_hh_t = whts[ci].data_ptr()
            ch.hsx = d["hsx"].data_ptr() + t0 * B * H * 2
            ch.gates = d["gates"].data_ptr() + t0 * _pad32(B) * H4 * 2
            ch.dg = dgs[ci].data_ptr() + t0 * B * H4 * 2
            if dhs[ci] is not None:
                ch.dhs = dhs[ci].data_ptr() + t0 * B * H * dhs[ci].element_size()
                ch.dhs_f32 = 0 if dhs[ci].dtype == BF16 else 1
            if last[ci] is not None:                        # gradient wrt the state this segment hands on
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
            LIB.call("fn_gru_seq_bwd_bf16", chains, n, B, L, H, _p(bar), bar.numel(), st)
            last = new_last
            if k < S:
                # gradient wrt cell 1's states of this segment = cell 2's input gradient: dgi W_ih2
                j = S - 1 - k
                t0 = j * L
                tc_gemm(dgs[3], t0 * B * H4, H4, 0, d2["w_ih_b"], 0, H, 1, dhs[2], t0 * B * H, H, None, L * B, H, K3)
                if j == 0:                                   # ... plus cell 2's initial-state gradient on hs_g[0]
                    LIB.call("fn_add_f32_to_bf16", _p(dhs[2]), _p(last[3]), B * H, st)

        out_grads = []
        for ci in range(3):
            out_grads += _finish_chain_grads(keep[ci]["spec"], keep[ci], dgs[ci], last[ci], B, T, H)
        out_grads += _finish_chain_grads(d2["spec"], d2, dgs[3], last[3], B, T, H, dxin_done=dhs[2])[:4]
        ctx.keep = None
        return (None, None, None, None) + tuple(out_grads)
