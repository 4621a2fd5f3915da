#!/usr/bin/env python
"""bench.py -- sequences/sec of one GM-VAE train step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1] [--impl reference]

One "step" = trainer_gmm.train(): zero_grad + forward + all losses + backward + global-norm clip
+ Adam (+ one NCCL all-reduce of the flat gradient buffer when N > 1) on one synthetic batch.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

# stdout carries exactly ONE JSON line.  Libraries write banners to the C-level stdout (NCCL / c10d print
# "NCCL version ..." when the first communicator is created), so file descriptor 1 is pointed at stderr for the whole
# run and the JSON line goes to a private duplicate of the original stdout.
if __name__ == "__main__":
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
else:                                   # imported: leave the importer's stdout alone
    _JSON_OUT = sys.stdout


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()

import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "music-fader-nets_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

WORKLOADS = {
    # name: (variant, B per GPU, T, H, Z, K, precision)
    "c1": ("vae", 4, 128, 256, 128, 0, "f32"),
    # BASELINE configs[1] ("fp32"): the fp32 parity bar (1e-3 relative, arg-max bit-exact) ON the tensor cores -- bf16x3 mode:
    # hi / lo bf16 operand planes, three plane products per product, fp32 accumulation; c2_f32 = the fp32 FMA (SIMT) kernels
    "c2": ("gmvae", 64, 256, 512, 128, 2, "bf16x3"),
    "c2_f32": ("gmvae", 64, 256, 512, 128, 2, "f32"),
    "c2_bf16": ("gmvae", 64, 256, 512, 128, 2, "bf16"),
    "c2_x3": ("gmvae", 64, 256, 512, 128, 2, "bf16x3"),
    "c3_x3": ("gmvae", 256, 512, 1024, 128, 2, "bf16x3"),
    "c3": ("gmvae", 256, 512, 1024, 128, 2, "bf16"),
    "c3_f32": ("gmvae", 256, 512, 1024, 128, 2, "f32"),
    # BASELINE configs[4]: arousal-transfer inference (encode -> shift z -> greedy decode), seq_len 512
    "c5": ("gmvae", 256, 512, 1024, 128, 2, "bf16"),
    "c5_f32": ("gmvae", 64, 512, 512, 128, 2, "f32"),
}
STEP0 = 20000          # beta0 = beta = 0.2: every KL term is live


def flops_per_token(H, V=342):
    """Algorithmic train-step FLOPs per (sequence, timestep): 3 x (54 H^2 + 2 H V)."""
    return 3.0 * (54.0 * H * H + 2.0 * H * V)


def gru_flops_per_token(H):
    """Recurrent-GEMM share handled by the persistent GRU kernels: 8 chains x 2*3H*H fwd, and the
    dgh*W_hh product of BPTT (same size) -> 2 x 48 H^2."""
    return 2.0 * 48.0 * H * H


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the GRU kernel launches of one step, from the committed
    `ncu --set full` capture of this workload (profiles/r02_traffic.json, else r01); None if it was not captured."""
    for tag in ("r02", "r01"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"))).get(workload)
        except Exception:
            v = None
        if v is not None:
            return v
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d.get("hbm_gbs"), tf_burst=d.get("bf16_tflops"), tf_sust=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batch_pinned(B, T, seed):
    """Synthetic event-token batch in PINNED host memory, in the dataset's tuple layout."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randint(2, 342, (B, T), generator=g)
    r = torch.randint(0, 3, (B, T), generator=g)
    n = torch.randint(0, 16, (B, T), generator=g)
    c = torch.rand(B, 24, generator=g)
    rd = (r == 1).double().mean(1)
    nd = n.double().mean(1)
    return [t.pin_memory() for t in (d, r, n, c, rd, nd)]


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import fadernets_b200 as fn
    from fadernets_b200 import trainer_gmm, trainer
    from fadernets_b200._lib import LIB

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    variant, B, T, H, Z, K, prec = WORKLOADS[args.workload]

    torch.manual_seed(0)
    if variant == "gmvae":
        model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K)
    else:
        model = fn.MusicAttrRegVAE(342, 3, 16, 24, H, Z, 32)
    model = model.to(dev).train()
    model.host_rng = True
    model.set_precision(prec)

    from fadernets_b200 import parallel
    if world > 1:
        model.flatten_parameters_()
        parallel.broadcast_parameters(model, src=0)
        torch.manual_seed(1000 + rank)          # every rank draws its own reparameterisation noise from here on
    # NCCL all-reduce of the flat gradient buffer in two buckets: the decoder-side bucket is issued from backward hooks
    # and runs under the encoder BPTT, the encoder bucket between backward and clip+Adam (parallel.OverlappedGradAllReduce)
    sync = None
    if world > 1:
        sync = parallel.GradAllReduce() if os.environ.get("FN_DP_OVERLAP", "1") == "0" else parallel.OverlappedGradAllReduce(model)
    opt = fn.FusedAdam(model, lr=1e-3, grad_sync=sync)
    tr = trainer_gmm if variant == "gmvae" else trainer
    if variant == "gmvae":
        tr.configure(model, opt, {"beta": 0.2, "lr": 1e-3})
    else:
        tr.configure(model, opt, {"beta": 0.2, "lr": 1e-3}, step_=STEP0)

    host = make_batch_pinned(B, T, seed=rank)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def device_inputs():
        d, r, n, c, rd, nd = [t.to(dev, non_blocking=True) for t in host]
        oh = [tr.convert_to_one_hot(x, k) for x, k in ((d, 342), (r, 3), (n, 16))]
        return oh, d, r, n, c, rd, nd

    def one_step(step, inputs):
        oh, d, r, n, c, rd, nd = inputs
        return tr.train(step, *oh, d, r, n, c, rd, nd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        """Returns (seconds for nsteps (max over ranks), launches)."""
        step = STEP0
        resident = None if e2e else device_inputs()
        barrier()
        l0 = LIB.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(nsteps):
            inputs = device_inputs() if e2e else resident
            step, out = one_step(step, inputs)          # `out` = 8 python floats: the D2H read of the result
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / 1e3, LIB.launches - l0, out

    # warm-up (also builds caches / allocator pools)
    timed(max(args.warmup, 3), e2e=False)
    sampler = ClockSampler(local).start() if rank == 0 else None
    sec, launches, out = timed(args.steps, e2e=False)
    clocks = sampler.stop() if sampler else None
    sec_e2e, _, _ = timed(args.steps, e2e=True)

    # ---- per-kernel time of the dominant kernel family (persistent GRU), measured live with CUDA events
    # on the launching stream; a separate short pass so the tracing does not perturb `value`.
    nprobe = min(args.steps, 3)
    rec, orig = [], LIB.call
    if rank == 0:
        def traced(name, *a):
            if name.startswith("fn_gru_seq_") or args.breakdown:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); rc = orig(name, *a); e1.record()
                if name == "fn_tc_gemm_bf16":
                    name = f"fn_tc_gemm_bf16 M{a[10]} N{a[11]} K{a[12]} amn{a[2]} bmn{a[5]} cbf{a[8]}"
                elif name == "fn_tc_gemm_bf16_splitk":
                    name = f"fn_tc_gemm_bf16_splitk M{a[10]} N{a[11]} K{a[12]} s{a[14]}"
                elif name == "fn_gemm_f32":
                    name = f"fn_gemm_f32 M{a[9]} N{a[10]} K{a[11]} sa{a[1]},{a[2]} sb{a[4]},{a[5]}"
                rec.append((name, e0, e1))
                return rc
            return orig(name, *a)
        LIB.call = traced
    timed(nprobe, e2e=False)                     # every rank runs it (collectives inside)
    LIB.call = orig
    gru_ms = sum(a.elapsed_time(b) for nm, a, b in rec if nm.startswith("fn_gru_seq_")) / nprobe if rec else None
    breakdown = None
    if args.breakdown and rec:
        agg = {}
        for nm, a, b in rec:
            t, c = agg.get(nm, (0.0, 0))
            agg[nm] = (t + a.elapsed_time(b), c + 1)
        breakdown = {k: [round(v[0] / nprobe, 3), v[1] // nprobe] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    gbatch = B * world
    value = gbatch * args.steps / sec
    e2e = gbatch * args.steps / sec_e2e
    tokens = B * T
    gru_flops = gru_flops_per_token(H) * tokens
    achieved = gru_flops / (gru_ms * 1e-3) / 1e12 if gru_ms else None
    line = {
        "metric": "sequences/sec GM-VAE train step", "value": round(value, 2), "unit": "sequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(sec / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": dict(workload_config(args.workload, world),
                       l2="per-step working set (saved gates + states, > 1 GB) exceeds the 126 MB L2; no flush needed",
                       weights="torch default init, manual_seed(0)", step=STEP0),
        "e2e": {"value": round(e2e, 2), "unit": "sequences/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 8 * 4, "ms_per_step": round(sec_e2e / args.steps * 1e3, 3)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": ("gru2_fwd_kernel + gru2_bwd_kernel (persistent tcgen05 cta_group::2 recurrent GEMM + gates, CTA pairs)" if prec == "bf16" else
                                "gru_tc_kernel<.., X3> (persistent tcgen05 recurrent GEMM over hi / lo bf16 planes: 3 tensor-core products per algorithmic product)" if prec == "bf16x3" else
                                "gru_fwd_kernel + gru_bwd_kernel (persistent fp32 SIMT recurrent GEMM + gates)"),
                     "bound": "tensor", "achieved": round(achieved, 3) if achieved else None,
                     "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                     "frac": round(achieved / peaks["tf_sust"], 5) if achieved else None,
                     "traffic": ncu_traffic(args.workload),
                     "peak_source": peaks["src"] + " cuBLAS bf16 sustained (kernel timed inside a long step)" +
                                    ("" if prec == "bf16" else "; achieved counts ALGORITHMIC FLOP (the tensor cores execute 3x that)" if prec == "bf16x3"
                                     else "; fp32 SIMT exact-parity path (FMA pipe, no tensor cores)"),
                     "flops_per_launch_group": gru_flops, "ms_per_step_in_kernel": round(gru_ms, 3) if gru_ms else None,
                     "step_flops": flops_per_token(H) * tokens,
                     "step_frac_of_peak": round(flops_per_token(H) * tokens * world / (sec / args.steps) / 1e12 / (peaks["tf_sust"] * world), 5)},
        "last_step_outputs": [round(float(x), 5) for x in out],
    }
    if breakdown:
        line["breakdown_ms_per_step"] = breakdown       # C-ABI call -> [ms per step, calls per step] (event-timed, serial)
    if world == 1 and not args.no_parity_mode and prec in ("bf16", "f32") and H % 64 == 0:
        # the SAME step in the mode that meets north_star's 1e-3 / bit-exact-arg-max bar on the tensor cores (bf16x3: hi / lo
        # bf16 operand planes, three plane products per product): what the parity tests of tests/test_gpu_parity.py run
        model.set_precision("bf16x3")
        timed(3, e2e=False)
        sec_p, _, out_p = timed(3, e2e=True)
        model.set_precision(prec)
        line["fp32_parity_mode"] = {"precision": "bf16x3", "ms_per_step": round(sec_p / 3 * 1e3, 3),
                                    "sequences_per_s": round(B * 3 / sec_p, 2), "steps": 3, "warmup": 3,
                                    "what": "e2e-timed train steps (host batches, H2D + result D2H inside) with every tensor-core "
                                            "product evaluated over hi/lo bf16 planes: fp32-level results (tests at rtol 1e-3)",
                                    "last_step_outputs": [round(float(x), 5) for x in out_p]}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference(args.workload, steps=3, warmup=1, budget_s=25.0)
    if not args.no_gpu_reference and world == 1:
        # the bar to beat: the reference's own GPU path -- cuDNN GRU, cuBLAS, ATen through torch -- on this B200
        from baseline import ref_runner
        why = ref_runner.available()
        if why is None:
            del model, opt
            torch.cuda.empty_cache()
            sampler = ClockSampler(local).start()
            gr = ref_runner.gpu_reference(variant, B, T, H, Z, K, steps=2, warmup=1)
            gr["clocks"] = sampler.stop()
            best = max((v.get("sequences_per_s", 0.0) for v in gr.values() if isinstance(v, dict)), default=0.0)
            gr["repo_e2e_over_reference_gpu"] = round(e2e / best, 2) if best else None
            if "fp32_parity_mode" in line and best:
                gr["repo_parity_mode_over_reference_gpu"] = round(line["fp32_parity_mode"]["sequences_per_s"] / best, 2)
            line["gpu_reference"] = gr
        else:
            line["gpu_reference"] = {"unavailable": why}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_decode(args):
    """BASELINE configs[4]: arousal transfer = encode(x) -> z + lambda * (mu_lookup(1) - mu_lookup(0)) -> eval-mode
    global_decoder(z, steps=T) (reference arousal_transfer.ipynb cells 11/15/17, test_class.py:233-254).  One
    "step" = one batch of B sequences; value = sequences/s (replicas only when N > 1: no collective)."""
    import fadernets_b200 as fn
    from fadernets_b200._lib import LIB
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    variant, B, T, H, Z, K, prec = WORKLOADS[args.workload]
    torch.manual_seed(0)
    model = fn.MusicAttrRegGMVAE(342, 3, 16, 24, H, Z, 32, n_component=K).to(dev).eval()
    model.set_precision(prec)
    host = make_batch_pinned(B, T, seed=rank)
    h2d = host[0].numel() * host[0].element_size() + host[3].numel() * host[3].element_size()

    def one(e2e):
        d = host[0].to(dev, non_blocking=True); c = host[3].to(dev, non_blocking=True)
        with torch.no_grad():
            dis_r, dis_n = model.encode(d)                       # token ids are accepted in place of the dense one-hot
            shift_r = model.mu_r_lookup.weight[1] - model.mu_r_lookup.weight[0]
            shift_n = model.mu_n_lookup.weight[1] - model.mu_n_lookup.weight[0]
            z = torch.cat([dis_r.mean + 0.5 * shift_r, dis_n.mean + 0.5 * shift_n, c], 1)
            _, toks = model.decode_greedy(z, T, return_logp=False)
        return toks.cpu() if e2e else toks

    def timed(n, e2e):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        l0 = LIB.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            toks = one(e2e)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / 1e3, LIB.launches - l0, toks

    timed(max(3, args.warmup), False)
    sampler = ClockSampler(local).start() if rank == 0 else None
    n = max(1, args.steps)
    sec, launches, toks = timed(n, False)
    clocks = sampler.stop() if sampler else None
    sec_e2e, _, _ = timed(min(n, 40), True)
    sec_e2e *= n / min(n, 40)
    # device time of the dominant kernel (the one-launch greedy decode), CUDA events around the C-ABI call
    rec, orig = [], LIB.call
    def traced(name, *a):
        if name == "fn_decode_greedy_bf16":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rc = orig(name, *a); e1.record(); rec.append((e0, e1))
            return rc
        return orig(name, *a)
    LIB.call = traced
    timed(3, False)
    LIB.call = orig
    dec_ms = sum(a.elapsed_time(b) for a, b in rec) / len(rec) if rec else None
    agreement = None
    if rank == 0 and prec == "bf16":
        # how often the bf16 tensor-core decode emits the token the fp32 exact-parity path emits (same weights, same z),
        # counted while both still follow the same prefix -- a reported number, not a threshold (64 sequences x 64 steps)
        with torch.no_grad():
            zs = torch.randn(64, 2 * Z + 24, generator=torch.Generator().manual_seed(5)).to(dev)
            _, t16 = model.decode_greedy(zs, 64, return_logp=False)
            model.set_precision("f32")
            _, t32 = model.decode_greedy(zs, 64, return_logp=False)
            model.set_precision(prec)
            same = (t16 == t32)
            agreement = {"tokens_equal": round(float(same.float().mean()), 4),
                         "tokens_equal_on_common_prefix": round(float(same.cumprod(1).float().mean()), 4),
                         "sample": "64 sequences x 64 steps, fp32 SIMT path as the reference"}
    if rank == 0:
        peaks = measured_peaks()
        fl = 2.0 * (18.0 * H * H + H * 342) * B * T              # cell 1 + cell 2 (input + recurrent) + projection
        line = {"metric": "sequences/sec greedy decode (arousal-transfer inference)", "value": round(B * world * n / sec, 2),
                "unit": "sequences/s", "n_gpus": world, "steps": n, "warmup": max(3, args.warmup),
                "ms_per_step": round(sec / n * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": prec, "data": "synthetic",
                "config": {"workload": f"{args.workload}: encode + latent shift + greedy global_decoder, {n * B} sequences per GPU in batches of "
                                       f"{B} x {T} steps, hidden {H}, {prec}", "global_batch": B * world, "sequences": n * B * world, "seq_len": T, "hidden": H,
                           "parallelism": f"replicas{world}", "l2": "weights (2 x 3H x H + H x V bf16) stay L2-resident by design"},
                "e2e": {"value": round(B * world * n / sec_e2e, 2), "unit": "sequences/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": B * T * 8},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "decode_tc_kernel (persistent greedy decode: cell 1 -> cell 2 -> projection -> arg-max, all steps in one launch)",
                             "bound": "tensor", "achieved": round(fl / (dec_ms * 1e-3) / 1e12, 3) if dec_ms else None, "peak": peaks["tf_sust"],
                             "unit": "TFLOP/s", "frac": round(fl / (dec_ms * 1e-3) / 1e12 / peaks["tf_sust"], 5) if dec_ms else None, "traffic": None,
                             "ms_per_launch": round(dec_ms, 3) if dec_ms else None, "flops_per_launch": fl,
                             "peak_source": peaks["src"] + " cuBLAS bf16 sustained"},
                "fp32_agreement": agreement,
                "tokens_head": toks[0, :8].tolist()}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_decode_reference(T, H, Z, K)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_decode_reference(T, H, Z, K, Bs=8):
    """The reference's own decode path on the host cores (arousal_transfer.ipynb cells 11/15: model.eval(); encode;
    shift; global_decoder(z, steps)) on a bounded sample: one batch of Bs sequences."""
    from baseline import ref_runner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    why = ref_runner.available()
    if why is not None:
        return {"unavailable": why}
    model, ns = ref_runner.make_trainer("gmvae", H, Z, K, "cpu")
    model.eval()
    d, r, n, c, rd, nd = ref_runner.synth_batch(Bs, T, 0, "cpu")
    d_oh = ns["convert_to_one_hot"](d, 342)
    with torch.no_grad():
        t0 = time.perf_counter()
        dis_r, dis_n = model.encode(d_oh)
        sr = model.mu_r_lookup.weight[1] - model.mu_r_lookup.weight[0]
        sn = model.mu_n_lookup.weight[1] - model.mu_n_lookup.weight[0]
        z = torch.cat([dis_r.mean + 0.5 * sr, dis_n.mean + 0.5 * sn, c], 1)
        out = model.global_decoder(z, steps=T)
        sec = time.perf_counter() - t0
    ref_runner._cpu_shim(False)
    return {"value": round(Bs / sec, 3), "unit": "sequences/s", "cores": cores, "kind": "reference",
            "sample": f"1 batch of {Bs} sequences x {T} steps, hidden {H}, fp32, unmodified reference classes in eval mode; {sec:.2f} s"}


SAMPLE_BATCH = {"c1": 4, "c2": 8, "c2_f32": 8, "c2_bf16": 8, "c2_x3": 8, "c3": 4, "c3_f32": 4, "c3_x3": 4}     # the batch at which the CPU reference is fastest per sequence


def workload_config(workload, world):
    """`config` of the JSON line -- shared by the repo arm and the reference arm so that they describe the same workload."""
    variant, B, T, H, Z, K, prec = WORKLOADS[workload]
    return {"workload": f"{workload}: Music{'AttrRegGMVAE' if variant == 'gmvae' else 'AttrRegVAE'} "
                        f"train step, batch {B}/GPU x seq_len {T}, hidden {H}, z {Z}, K {K}, {prec}",
            "global_batch": B * world, "seq_len": T, "hidden": H, "parallelism": f"dp{world}"}


def cpu_reference(workload, steps, warmup, budget_s=None):
    """The reference's own CPU implementation on the host cores: the UNMODIFIED reference classes and its own
    train() (baseline/_ref, see baseline/ref_runner.py) on a BOUNDED sample of the workload (reduced batch, same
    seq_len / hidden; CPU sequences/s is ~flat in batch).  Falls back to the oracle port
    (kind "port") only if baseline/_ref was not vendored."""
    variant, B, T, H, Z, K, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = SAMPLE_BATCH[workload]
    from baseline import ref_runner
    why = ref_runner.available()
    if why is None:
        sec, done, last = ref_runner.time_train(variant, Bs, T, H, Z, K, "cpu", steps, warmup, budget_s=budget_s)
        kind = "reference"
    else:
        from oracle import fader_oracle as fo
        w = fo.init_weights(H, Z, variant, max(K, 1), seed=0)
        st = fo.AdamState(w)
        batch = fo.synth_batch(Bs, T, seed=0)
        g = torch.Generator().manual_seed(1)
        times = []
        for it in range(warmup + steps):
            er, en = torch.randn(Bs, Z, generator=g), torch.randn(Bs, Z, generator=g)
            t0 = time.perf_counter()
            fo.train_step(w, st, variant, batch, er, en, STEP0, 0.2, 1e-3, dense_onehot=True)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        sec, done, kind = float(np.mean(times)), steps, "port"
    return {"value": round(Bs / sec, 4), "unit": "sequences/s", "cores": cores, "kind": kind,
            "sample": f"{done} timed + {warmup} warm-up train step(s) of batch {Bs} (of {B}) x seq_len {T}, hidden {H}, fp32; "
                      f"{sec:.2f} s/step; CPU sequences/s is ~flat in batch (BASELINE.md)"
                      + ("" if why is None else f"; oracle port because {why}"),
            "sample_batch": Bs, "steps": done, "warmup": warmup, "ms_per_step": round(sec * 1e3, 1)}


def run_reference(args):
    """--impl reference: the unmodified reference's CPU train step on all host cores.  Exactly `--warmup` + `--steps`
    steps run; one step = one train() call on a bounded sample (SAMPLE_BATCH sequences of the workload's seq_len /
    hidden), so value = sample_batch / seconds per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    variant, B, T, H, Z, K, prec = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = cpu_reference(args.workload, steps=max(1, args.steps), warmup=max(0, args.warmup))
    cfg = workload_config(args.workload, world)
    cfg.update({"sample_batch": cb["sample_batch"],
                "note": "unmodified reference classes + train() (fp32, nn.GRU on dense one-hots) on the host cores; each step = "
                        "one train() on a bounded sample of the workload (sample_batch sequences, same seq_len / hidden)"})
    line = {"impl": "reference", "metric": "sequences/sec GM-VAE train step", "value": cb["value"],
            "unit": "sequences/s", "n_gpus": world, "steps": cb["steps"], "warmup": cb["warmup"],
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 10; c5: 391 batches of 256 = 100k sequences)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="c3 (default) = BASELINE configs[2]: the per-GPU shape of the 1/2/4/8-GPU metric (configs[3] = 8 x c3); "
                         "c2 = configs[1] (fp32 exact-parity path); c1 = configs[0]")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the unmodified reference on this GPU")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the extra bf16x3 (fp32 parity on the tensor cores) timing")
    ap.add_argument("--breakdown", action="store_true", help="add per-C-ABI-call device time to the JSON line")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 391 if args.workload == "c5" else 98 if args.workload == "c5_f32" else 10
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
        (run_decode if args.workload.startswith("c5") else run_ours)(args)


if __name__ == "__main__":
    main()
