"""HBM-bound kernels of the path against the measured copy bandwidth: vocabulary log-softmax / NLL,
clip + Adam, gradient norm, and the GM latent block (reparameterised sampling, responsibilities, mixture KL) at the
BASELINE shape (B = 256 sequences: 1 MB, launch-latency bound) and at an inflated row count (2^20 sequences) where
bandwidth is what is measured.  Prints one JSON object; `python tools/bw_bench.py > profiles/rNN_bandwidth.json`.
Algorithmic bytes only (each operand read or written once)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "music-fader-nets_b200")); sys.path.insert(0, ROOT)
import torch
from fadernets_b200._lib import LIB
from fadernets_b200.ops import _p, _st

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK = float(peaks.get("hbm_gbs", 6550.0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2


def timeit(fn, bytes_, reps=10):
    for _ in range(3): fn()
    ms = []
    for _ in range(reps):
        flush.zero_()                                               # evict the operands from L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    t = ms[len(ms) // 2]
    return {"ms": round(t, 4), "algorithmic_bytes": int(bytes_), "GB/s": round(bytes_ / t / 1e6, 1), "frac_of_measured_peak": round(bytes_ / t / 1e6 / PEAK, 3)}


out = {"peak_GB/s": PEAK, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peaks else "fallback 6550", "l2": "256 MB flush between launches", "kernels": {}}
K = out["kernels"]
g = torch.Generator(device=dev).manual_seed(0)

# ---- vocabulary head, config 3: B=256, T=512, V=342 (fp32 logits time-major) ------------------------------------
B, T, V = 256, 512, 342
logits = torch.randn(T, B, V, device=dev, generator=g)
target = torch.randint(0, V, (B, T), device=dev, generator=g)
outp = torch.empty(B, T, V, device=dev); lse = torch.empty(T * B, device=dev); rows = torch.empty(T * B, device=dev)
dl = torch.empty(T, B, V, device=dev); dout = torch.randn(B, T, V, device=dev, generator=g)
one = torch.ones(1, device=dev)
nb = T * B * V * 4
K["vocab_nll_fwd (log-softmax + NLL rows, writes log-probs)"] = timeit(lambda: LIB.call("fn_vocab_nll_fwd", _p(logits), _p(target), B, T, V, _p(outp), _p(lse), _p(rows), _st(lse)), 2 * nb)
K["vocab_nll_fwd (loss only, no log-prob output)"] = timeit(lambda: LIB.call("fn_vocab_nll_fwd", _p(logits), _p(target), B, T, V, None, _p(lse), _p(rows), _st(lse)), nb)
K["vocab_nll_bwd (softmax - onehot)"] = timeit(lambda: LIB.call("fn_vocab_nll_bwd", _p(logits), _p(lse), _p(target), _p(one), 1.0 / (T * B), B, T, V, _p(dl), _st(dl)), 2 * nb)
K["vocab_logsoftmax_fwd"] = timeit(lambda: LIB.call("fn_vocab_logsoftmax_fwd", _p(logits), B, T, V, _p(outp), _st(outp)), 2 * nb)
K["vocab_logsoftmax_bwd"] = timeit(lambda: LIB.call("fn_vocab_logsoftmax_bwd", _p(outp), _p(dout), B, T, V, _p(dl), _st(dl)), 3 * nb)
del logits, outp, dl, dout

# ---- optimiser, H=1024: 37.29 M live parameters ------------------------------------------------------------------
n = 37_290_000
p = torch.randn(n, device=dev, generator=g); gr = torch.randn(n, device=dev, generator=g) * 1e-3
m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); norm = torch.empty(1, device=dev)
sb = LIB.call("fn_reduce_scratch_bytes", n); scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=dev)
K["grad_norm"] = timeit(lambda: LIB.call("fn_grad_norm", _p(gr), n, _p(norm), _p(scratch), sb, _st(gr)), 4 * n)
K["clip_adam"] = timeit(lambda: LIB.call("fn_clip_adam", _p(p), _p(gr), _p(m), _p(v), n, _p(norm), 1.0, 1e-4, 0.9, 0.999, 1e-8, 1, _st(p)), 28 * n)
del p, gr, m, v

# ---- latent block: Z=128, K=2; BASELINE shape and inflated ------------------------------------------------------------
Z, KC = 128, 2
for Bn, tag in ((256, "B=256 (config 3)"), (1 << 20, "B=2^20 (inflated)")):
    mu = torch.randn(Bn, Z, device=dev, generator=g); pre = torch.randn(Bn, Z, device=dev, generator=g) * 0.1
    eps = torch.randn(Bn, Z, device=dev, generator=g); scale = torch.empty(Bn, Z, device=dev); z = torch.empty(Bn, Z, device=dev)
    mul = torch.randn(KC, Z, device=dev, generator=g); lvl = torch.full((KC, Z), -4.0, device=dev)
    ll = torch.empty(Bn, KC, device=dev); qy = torch.empty(Bn, KC, device=dev); y = torch.empty(Bn, dtype=torch.int64, device=dev)
    out3 = torch.empty(3, device=dev)
    wsb = LIB.call("fn_latent_scratch_bytes", Bn, Z, KC); ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    nz = Bn * Z * 4
    K[f"reparam_fwd {tag}"] = timeit(lambda: LIB.call("fn_reparam_fwd", _p(mu), _p(pre), _p(eps), Bn * Z, _p(scale), _p(z), _st(z)), 5 * nz)
    K[f"qy_x_fwd {tag}"] = timeit(lambda: LIB.call("fn_qy_x_fwd", _p(z), _p(mul), _p(lvl), Bn, Z, KC, _p(ll), _p(qy), _p(y), _st(z)), nz + Bn * (2 * KC * 4 + 8))
    K[f"gm_kl_fwd {tag}"] = timeit(lambda: LIB.call("fn_gm_kl_fwd", _p(mu), _p(scale), _p(mul), _p(lvl), _p(qy), _p(ll), None, 0, Bn, Z, KC, _p(out3), _p(ws), wsb, _st(mu)), 2 * nz + Bn * 2 * KC * 4)
    dmu = torch.empty(Bn, Z, device=dev); dsc = torch.empty(Bn, Z, device=dev); dqy = torch.empty(Bn, KC, device=dev); dll = torch.empty(Bn, KC, device=dev)
    dmul = torch.zeros(KC, Z, device=dev); d3 = torch.ones(3, device=dev)
    K[f"gm_kl_bwd {tag}"] = timeit(lambda: LIB.call("fn_gm_kl_bwd", _p(mu), _p(scale), _p(mul), _p(lvl), _p(qy), _p(ll), None, 0, Bn, Z, KC, _p(d3), _p(dmu), _p(dsc), _p(dqy), _p(dll), _p(dmul), _p(ws), wsb, _st(mu)), 4 * nz + Bn * 4 * KC * 4)
    K[f"std_kl_fwd {tag}"] = timeit(lambda: LIB.call("fn_std_kl_fwd", _p(mu), _p(scale), Bn * Z, _p(out3), _p(ws), wsb, _st(mu)), 2 * nz)
    del mu, pre, eps, scale, z, dmu, dsc
print(json.dumps(out, indent=1))
