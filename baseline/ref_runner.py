"""Drives the UNMODIFIED reference (gudgud96/music-fader-nets) for bench.py's reference arm and its
`gpu_reference` field.  BENCH INFRASTRUCTURE: nothing under music-fader-nets_b200/ imports this.

The reference is a flat script directory with no setup.py / pyproject, so `pip install --target baseline/_ref
/root/reference` has nothing to install; instead `__graft_entry__.build()` copies the four files of the
hot path (gmm_model.py, model_v2.py, trainer_gmm.py, trainer.py) verbatim into the git-ignored
`baseline/_ref/` (it travels to the GPU box with the snapshot; /root/reference does not exist there).
How they are driven (same recipe as oracle/gen_golden.py):
  * the two model files import cleanly (torch + numpy only) and are used as they are;
  * the trainer scripts train at import time and need absent packages, so their step functions
    (std_normal, loss_function, latent_regularized_loss_function, train, evaluate, convert_to_one_hot) are
    AST-extracted and exec'd, unmodified, in a namespace holding `model, optimizer, args, step`;
  * on a CPU-only run the hard-coded `.cuda()` calls are neutralised (`torch.Tensor.cuda = identity`,
    `nn.Module.cuda = identity`); on the GPU they run as written.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_FILES = ("gmm_model.py", "model_v2.py", "trainer_gmm.py", "trainer.py")
STEP_FUNCS = ("std_normal", "loss_function", "latent_regularized_loss_function", "train", "evaluate",
              "convert_to_one_hot")
STEP0 = 20000            # beta0 = beta: every KL term live (same as the repo arm)


def available() -> str | None:
    """None when the vendored reference files are present, else the reason."""
    missing = [f for f in REF_FILES if not os.path.exists(os.path.join(REF_DIR, f))]
    return f"baseline/_ref misses {missing} (run __graft_entry__.build() where /root/reference exists)" if missing else None


def _load_module(name):
    spec = importlib.util.spec_from_file_location("fader_ref_" + name, os.path.join(REF_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_CUDA_SHIM = {}


def _cpu_shim(on: bool):
    """CPU runs only: the reference calls .cuda() unconditionally (gmm_model.py:120,212,214,230)."""
    if on and not _CUDA_SHIM:
        _CUDA_SHIM["t"], _CUDA_SHIM["m"] = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    elif not on and _CUDA_SHIM:
        torch.Tensor.cuda, torch.nn.Module.cuda = _CUDA_SHIM.pop("t"), _CUDA_SHIM.pop("m")


def make_trainer(variant: str, H: int, Z: int, K: int, device: str, lr=1e-3, beta=0.2, seed=0):
    """(model, namespace with the reference's own train()/loss functions) on `device`."""
    from torch import nn, optim
    from torch.distributions import Normal, kl_divergence
    from torch.nn import functional as F
    _cpu_shim(device == "cpu")
    torch.manual_seed(seed)
    if variant == "gmvae":
        model = _load_module("gmm_model").MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
        trainer = "trainer_gmm.py"
    else:
        model = _load_module("model_v2").MusicAttrRegVAE(342, 3, 16, 24, H, Z, 32)
        trainer = "trainer.py"
    if device != "cpu":
        model.cuda()
    model.train()
    args = dict(lr=lr, beta=beta)
    optimizer = optim.Adam(model.parameters(), lr=lr)
    ns = dict(model=model, optimizer=optimizer, args=args, step=STEP0, np=np, torch=torch, F=F, nn=nn, optim=optim,
              Normal=Normal, kl_divergence=kl_divergence)
    tree = ast.parse(open(os.path.join(REF_DIR, trainer)).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in STEP_FUNCS]
    exec(compile(ast.Module(body=body, type_ignores=[]), trainer, "exec"), ns)
    return model, ns


def synth_batch(B, T, seed, device):
    """Same synthetic recipe as bench.make_batch_pinned (ids, attribute ids, chroma, densities)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randint(2, 342, (B, T), generator=g)
    r = torch.randint(0, 3, (B, T), generator=g)
    n = torch.randint(0, 16, (B, T), generator=g)
    c = torch.rand(B, 24, generator=g)
    rd = (r == 1).double().mean(1).numpy()
    nd = n.double().mean(1).numpy()
    dev = torch.device(device)
    return d.to(dev), r.to(dev), n.to(dev), c.to(dev), rd, nd


def time_train(variant, B, T, H, Z, K, device, steps, warmup, budget_s=None):
    """Runs the reference's own train() `warmup` + `steps` times on one synthetic batch of B sequences.
    Returns (seconds per timed step, timed steps actually run, last outputs).  `budget_s` bounds the timed
    loop (at least one timed step always runs)."""
    model, ns = make_trainer(variant, H, Z, K, device)
    d, r, n, c, rd, nd = synth_batch(B, T, 0, device)
    oh = [ns["convert_to_one_hot"](x, k) for x, k in ((d, 342), (r, 3), (n, 16))]
    cuda = device != "cpu"

    def one(step):
        ns["step"] = step                                   # trainer.py reads the module global
        return ns["train"](step, *oh, d, r, n, c, rd, nd)

    step = STEP0
    out = None
    for _ in range(warmup):
        step, out = one(step)
    if cuda:
        torch.cuda.synchronize()
    done, t0 = 0, time.perf_counter()
    if cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for _ in range(steps):
        step, out = one(step)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
    else:
        sec = time.perf_counter() - t0
    _cpu_shim(False)
    return sec / done, done, [float(x) for x in out]


def gpu_reference(variant, B, T, H, Z, K, steps=2, warmup=1):
    """The reference's own GPU path (cuDNN GRU + cuBLAS + ATen through torch) on cuda:current, full shape:
    TF32 as torch ships it (cudnn.allow_tf32 True, matmul TF32 off) and strict fp32."""
    out = {}
    for tag, tf32 in (("default_flags", True), ("allow_tf32_false", False)):
        old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        try:
            torch.backends.cudnn.allow_tf32 = tf32
            if not tf32:
                torch.backends.cuda.matmul.allow_tf32 = False
            sec, done, last = time_train(variant, B, T, H, Z, K, "cuda", steps, warmup)
            out[tag] = {"ms_per_step": round(sec * 1e3, 2), "sequences_per_s": round(B / sec, 2), "steps": done,
                        "warmup": warmup, "loss": round(last[0], 4)}
        except Exception as e:                      # e.g. out of memory: report, never take the repo arm down
            out[tag] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
            torch.cuda.empty_cache()
    out["what"] = (f"unmodified reference classes + its own train() (baseline/_ref) on this GPU, batch {B} x seq_len {T}, "
                   f"hidden {H}, fp32 parameters; torch {torch.__version__}")
    return out
