"""Profiling aid: throughput of fn_tc_gemm_bf16(_splitk) on the train step's batched-product shapes (config 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200")); sys.path.insert(0, ROOT)
import torch
from fadernets_b200._lib import LIB
from fadernets_b200.ops import _p, _st
from fadernets_b200.ops_bf16 import plan_splits

dev = torch.device("cuda:0")
bf = torch.bfloat16
TB, H = 131072, 1024
# (name, M, N, K, a_mn, b_mn, c_bf16, splits)
shapes = [("fwd  x W^T (segment)", TB // 16, 3 * H, H, 0, 0, 1, 1),
          ("dgrad dy W (segment)", TB // 16, H, 3 * H, 0, 1, 1, 1),
          ("wgrad dW_hh[:2H] splitK", 2 * H, H, TB, 1, 1, 0, 2),
          ("wgrad dW_hh[2H:] splitK", H, H, TB, 1, 1, 0, 4),
          ("wgrad dW_ih onehot splitK", 3 * H, 342, TB, 1, 1, 0, 4),
          ("logits h W_out^T   ", TB, 342, H, 0, 0, 0, 1),
          ("dgrad logits dy W_out", TB, H, 342, 0, 1, 1, 1),
          ("square 8192^3      ", 8192, 8192, 8192, 0, 0, 1, 1)]
for name, M, N, K, a_mn, b_mn, c_bf16, splits in shapes:
    splits = plan_splits(M, N, K)              # what the train step uses (FN_GEMM_PAIR=0: the 128x128-tile kernel only)
    r8 = lambda x: (x + 7) // 8 * 8
    def padded(rows, cols):                      # TMA row pitch: a multiple of 16 bytes
        buf = torch.zeros(rows, r8(cols), device=dev, dtype=bf)
        buf[:, :cols] = (torch.randn(rows, cols, device=dev) * 0.05).to(bf)
        return buf[:, :cols]
    A = padded(K, M) if a_mn else padded(M, K)
    B = padded(K, N) if b_mn else padded(N, K)
    ldc = (N + 7) // 8 * 8
    C = torch.zeros(M, ldc, device=dev, dtype=bf if c_bf16 else torch.float32)
    ws_bytes = LIB.call("fn_tc_gemm_splitk_ws_bytes", M, N, splits) if splits > 1 else 0
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    def run():
        if splits > 1:
            LIB.call("fn_tc_gemm_bf16_splitk", _p(A), A.stride(0), a_mn, _p(B), B.stride(0), b_mn, _p(C), ldc, c_bf16, None, M, N, K, 0, splits, _p(ws), ws_bytes, _st(C))
        else:
            LIB.call("fn_tc_gemm_bf16", _p(A), A.stride(0), a_mn, _p(B), B.stride(0), b_mn, _p(C), ldc, c_bf16, None, M, N, K, 0, _st(C))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # correctness spot check against torch on a corner
    Af = (A.t() if a_mn else A)[:256].float(); Bf = (B if b_mn else B.t())[:, :256].float()
    ref = Af @ Bf
    err = (C[:256, :256].float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    # the library GEMM on the same operands (cuBLAS through torch.matmul), same output dtype
    At = (A.t() if a_mn else A); Bt = (B if b_mn else B.t())
    for _ in range(3): torch.matmul(At, Bt)
    e0.record()
    for _ in range(10): torch.matmul(At, Bt)
    e1.record(); torch.cuda.synchronize()
    ms_cb = e0.elapsed_time(e1) / 10
    print(f"{name} M={M} N={N} K={K} splits={splits}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  rel.err {err:.2e}   | cuBLAS {ms_cb:.3f} ms {2.0 * M * N * K / ms_cb / 1e9:.0f} TFLOP/s  ratio {ms_cb / ms:.2f}")
