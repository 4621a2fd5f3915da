"""Run-to-run determinism probe (GPU): repeats one c3-sized bf16 train step and reports which gradient tensors / outputs differ
between repetitions (there are no float atomics and every split reduction has a fixed order, so nothing may differ)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "music-fader-nets_b200"), os.path.join(ROOT, "tests")]
import torch
import fadernets_b200 as fn
from test_gpu_fullsize import _batch, _step

dev = torch.device("cuda:0")
B, T, H, Z = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 512, 1024, 128)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 8
prec = sys.argv[6] if len(sys.argv) > 6 else "bf16"
torch.manual_seed(0)
model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=2).to(dev).train().set_precision(prec)
batch = _batch(B, T, 1, dev)
g = torch.Generator().manual_seed(2)
eps = [torch.randn(B, Z, generator=g).to(dev) for _ in range(2)]
ref = None
bad = 0
for r in range(reps):
    loss, out, r_out, n_out = _step(model, batch, eps)
    cur = {k: p.grad.detach().clone() for k, p in model.live_parameters()}
    cur["__out"], cur["__loss"], cur["__r_out"], cur["__n_out"] = out.clone(), loss.clone(), r_out.clone(), n_out.clone()
    if ref is None:
        ref = cur
        continue
    diff = [(k, float((cur[k].float() - ref[k].float()).abs().max()), int((cur[k] != ref[k]).sum()), cur[k].numel()) for k in cur if not torch.equal(cur[k], ref[k])]
    if diff:
        bad += 1
        print(f"rep {r} vs rep 0: {len(diff)} tensors differ (name, max abs, #elements, of):", sorted(diff, key=lambda kv: -kv[1])[:12], flush=True)
print(f"{prec} B={B} T={T} H={H}: {bad} of {reps - 1} repetitions differed from the first")
