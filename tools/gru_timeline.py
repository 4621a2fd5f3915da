"""Profiling aid: per-step pipeline timeline of CTA 0 of the tcgen05 GRU forward kernel (clock64 stamps).
usage: python tools/gru_timeline.py B T H n_chains"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200")); sys.path.insert(0, ROOT)
import math
import torch
from fadernets_b200._lib import LIB
from fadernets_b200.ops import ChainSpec, _p
from fadernets_b200.ops_bf16 import GruGroupBf16Fn

B, T, H, NCH = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda:0")
V = 342
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, V, (T, B), generator=g).int().to(dev)
specs, tensors = [], []
for c in range(NCH):
    specs.append(ChainSpec(emb_cols=(0, V), ids=ids, reverse=bool(c & 1), final=(0, c * H)))
    k = 1 / math.sqrt(H)
    tensors += [(torch.randn(3 * H, V, generator=g) * k).to(dev).requires_grad_(True), (torch.randn(3 * H, generator=g) * k).to(dev).requires_grad_(True),
                (torch.randn(3 * H, H, generator=g) * k).to(dev).requires_grad_(True), (torch.randn(3 * H, generator=g) * k).to(dev).requires_grad_(True)]
dbg = torch.zeros((T + 1) * 2 * 64, dtype=torch.int64, device=dev)
for it in range(2):
    (fin,) = GruGroupBf16Fn.apply(specs, B, T, H, (NCH * H,), *tensors)
torch.cuda.synchronize()
LIB.call("fn_gru_debug_timeline", _p(dbg))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
(fin,) = GruGroupBf16Fn.apply(specs, B, T, H, (NCH * H,), *tensors)
e1.record()
torch.cuda.synchronize()
LIB.call("fn_gru_debug_timeline", None)
print(f"forward launch (incl. setup kernels): {e0.elapsed_time(e1):.3f} ms = {e0.elapsed_time(e1) / T * 1e3:.2f} us/step")
d = dbg.view(T + 1, 2, 64).cpu()
names = ["prod:start", "prod:flag seen", "prod:fenced", "prod:tma issued", "mma:first chunk", "mma:committed", "epi:arrive", "epi:acc ready",
         "epi:math done", "epi:stores issued", "epi:published"]
nbt = (B + 127) // 128
for s in range(T // 2, min(T // 2 + 3, T)):
    for bt in range(nbt):
        t0 = int(d[s, bt, 0])
        print(f"step {s} bt {bt}: " + "  ".join(f"{names[k]}={int(d[s, bt, k]) - t0}" for k in range(1, 11)))
        print(f"      mma warp: turn starts={int(d[s, bt, 11]) - t0} acc_empty seen={int(d[s, bt, 12]) - t0}  per stage (weights ready, state ready): "
              + " ".join(f"({int(d[s, bt, 16 + 2 * j]) - t0},{int(d[s, bt, 17 + 2 * j]) - t0})" for j in range(min(12, H // 128))))
    if s + 1 < T:
        print(f"   step period (prod:start to next prod:start, bt0): {int(d[s + 1, 0, 0]) - int(d[s, 0, 0])} cycles")
# backward timeline too
loss = fin.sum()
dbg.zero_()
LIB.call("fn_gru_debug_timeline", _p(dbg))
e0.record(); loss.backward(); e1.record(); torch.cuda.synchronize()
LIB.call("fn_gru_debug_timeline", None)
print(f"backward (incl. weight-gradient GEMMs): {e0.elapsed_time(e1):.3f} ms")
d = dbg.view(T + 1, 2, 64).cpu()
for i in range(T // 2, min(T // 2 + 2, T)):
    for bt in range(nbt):
        t0 = int(d[i, bt, 0])
        print(f"bwd iter {i} bt {bt}: " + "  ".join(f"{names[k]}={int(d[i, bt, k]) - t0}" for k in (1, 2, 3, 4, 5, 6, 7, 9, 10)))
    print(f"   iter period: {int(d[i + 1, 0, 0]) - int(d[i, 0, 0])} cycles")
