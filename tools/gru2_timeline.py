"""Profiling aid: per-step pipeline timeline of CTA 0 of the CTA-pair GRU forward kernel (clock64 stamps, fn_gru_tc2.cu).
usage: python tools/gru2_timeline.py B T H n_chains"""
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
dbg = torch.zeros((T + 1) * 64, dtype=torch.int64, device=dev)
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
d = dbg.view(T + 1, 64).cpu()
names = {1: "ld:flag", 2: "ld:issued", 3: "mma:first data", 4: "mma:last commit", 5: "epi:x loaded", 6: "epi:acc ready",
         9: "st:stored", 10: "st:published", 11: "epi:gates stored"}
for s in range(T // 2, min(T // 2 + 3, T - 1)):
    t0 = int(d[s, 0])
    for k in range(2):
        o = k * 32
        if not int(d[s, o + 1]):
            continue
        print(f"step {s} chain {k}: ld:start={int(d[s, o]) - t0}  " + "  ".join(f"{names[e]}={int(d[s, o + e]) - t0}" for e in sorted(names)))
        print("      epi warps staged: " + " ".join(str(int(d[s, o + 12 + j]) - t0) for j in range(16)))
    print(f"   step period: {int(d[s + 1, 0]) - t0} cycles")
