"""Profiling aid: one launch each of the streaming latent kernels at B = 2^20 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200")); sys.path.insert(0, ROOT)
import torch
from fadernets_b200._lib import LIB
from fadernets_b200.ops import _p, _st
dev = torch.device("cuda:0")
Bn, Z, KC = 1 << 20, 128, 2
g = torch.Generator(device=dev).manual_seed(0)
mu = torch.randn(Bn, Z, device=dev, generator=g); scale = torch.rand(Bn, Z, device=dev, generator=g) + 0.5
mul = torch.randn(KC, Z, device=dev, generator=g); lvl = torch.full((KC, Z), -4.0, device=dev)
ll = torch.empty(Bn, KC, device=dev); qy = torch.empty(Bn, KC, device=dev); y = torch.empty(Bn, dtype=torch.int64, device=dev)
out3 = torch.empty(3, device=dev)
wsb = LIB.call("fn_latent_scratch_bytes", Bn, Z, KC); ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
for _ in range(2):
    LIB.call("fn_qy_x_fwd", _p(mu), _p(mul), _p(lvl), Bn, Z, KC, _p(ll), _p(qy), _p(y), _st(mu))
    LIB.call("fn_gm_kl_fwd", _p(mu), _p(scale), _p(mul), _p(lvl), _p(qy), _p(ll), None, 0, Bn, Z, KC, _p(out3), _p(ws), wsb, _st(mu))
torch.cuda.synchronize()
print(out3.tolist())
