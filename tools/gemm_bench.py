"""Profiling aid: throughput of fn_tc_gemm_bf16(_splitk) on the train step's batched-product shapes (config 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200")); sys.path.insert(0, ROOT)
import torch
from fadernets_b200._lib import LIB
from fadernets_b200.ops import _p, _st

dev = torch.device("cuda:0")
bf = torch.bfloat16
TB, H = 131072, 1024
# (name, M, N, K, a_mn, b_mn, c_bf16, splits)
shapes = [("fwd  x W^T        ", TB, 3 * H, H, 0, 0, 1, 1),
          ("dgrad dy W         ", TB, H, 3 * H, 0, 1, 1, 1),
          ("wgrad dy^T x splitK", 3 * H, H, TB, 1, 1, 0, 8),
          ("logits h W_out^T   ", TB, 342, H, 0, 0, 0, 1),
          ("square 8192^3      ", 8192, 8192, 8192, 0, 0, 1, 1)]
for name, M, N, K, a_mn, b_mn, c_bf16, splits in shapes:
    A = (torch.randn((K, M) if a_mn else (M, K), device=dev) * 0.05).to(bf)
    B = (torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.05).to(bf)
    ldc = (N + 7) // 8 * 8
    C = torch.zeros(M, ldc, device=dev, dtype=bf if c_bf16 else torch.float32)
    ws_bytes = LIB.call("fn_tc_gemm_splitk_ws_bytes", M, N, splits) if splits > 1 else 0
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    def run():
        if splits > 1:
            LIB.call("fn_tc_gemm_bf16_splitk", _p(A), A.shape[1], a_mn, _p(B), B.shape[1], b_mn, _p(C), ldc, c_bf16, None, M, N, K, 0, splits, _p(ws), ws_bytes, _st(C))
        else:
            LIB.call("fn_tc_gemm_bf16", _p(A), A.shape[1], a_mn, _p(B), B.shape[1], b_mn, _p(C), ldc, c_bf16, None, M, N, K, 0, _st(C))
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
    print(f"{name} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  rel.err {err:.2e}")
