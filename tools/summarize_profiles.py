"""Turns the raw ncu outputs of tools/profile.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py rNN"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list (ncu --metrics gpu__time_duration.sum --csv): per-kernel totals of the LAST train step ---------------
src = os.path.join(G, f"{tag}_launches_c3_bf16.csv")
if os.path.exists(src):
    lines = [l for l in open(src, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    recs = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3))
            for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    # one step = the launches after the last clip_adam of the warm-up step
    idx = [i for i, (n, _) in enumerate(recs) if "clip_adam" in n]
    step = recs[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else recs
    tot = sum(t for _, t in step)
    agg = {}
    for n, t in step:
        a = agg.setdefault(n, [0.0, 0]); a[0] += t; a[1] += 1
    with open(os.path.join(P, f"{tag}_launches_c3_bf16.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline\n")
        f.write(f"# {tag}, bf16 tensor-core path, B=256 T=512 H=1024: the {len(step)} launches of ONE train step (between two clip_adam launches);\n")
        f.write("# per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes\n")
        f.write(f"# total {tot:.1f} us\n")
        for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{t:12.1f} us {100 * t / tot:6.2f}% n={c:4d} {n[:150]}\n")
    with open(os.path.join(P, f"{tag}_launches_c3_bf16.csv"), "w") as f:
        f.write("kernel,us\n")
        for n, t in step: f.write(f'"{n}",{t:.3f}\n')
    print("launch list:", len(step), "launches,", round(tot / 1e3, 2), "ms")

# ---- full capture of the GRU kernels -------------------------------------------------------------------------------
rep = os.path.join(G, f"{tag}_gru_tc_c3.ncu-rep")
raw = os.path.join(G, f"{tag}_gru_tc_c3_raw.csv")          # `ncu -i ... --page raw --csv` already run on the GPU box (the report itself is too big to bring back)
if os.path.exists(rep) or os.path.exists(raw):
    out = open(raw).read() if os.path.exists(raw) else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    keep = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    traffic = 0.0
    with open(os.path.join(P, f"{tag}_ncu_gru_tc_c3_bf16.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, bench.py --workload c3 (B=256,T=512,H=1024,bf16): the {len(rows) - 2} GRU kernel launches of ONE train step\n")
        f.write("# (FN_GRU2_COOP=0: ncu cannot replay cooperative cluster launches; same kernels, same cluster shape)\n")
        tens = {}
        for r in rows[2:]:
            nm = r[h.index('Kernel Name')]
            kk = "fwd" if "fwd" in nm else "bwd"
            try:
                dur = float(r[h.index("gpu__time_duration.sum")]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(units[h.index("gpu__time_duration.sum")], 1.0)
                tp = float(r[h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")])
                a = tens.setdefault(kk, [0.0, 0.0]); a[0] += dur * tp; a[1] += dur
            except Exception:
                pass
        for kk, (num, den) in tens.items():
            f.write(f"# {kk}: duration-weighted sm__pipe_tensor_cycles_active (pct of peak, elapsed) = {num / max(den, 1e-9):.1f} % over {den / 1e3:.2f} ms of launches\n")
        for r in rows[2:]:
            f.write(f"--- {r[h.index('Kernel Name')][:110]}\n")
            for k in keep:
                if k in h: f.write(f"{k:84s} {r[h.index(k)]:>16s} {units[h.index(k)]}\n")
            def gb(k):
                v, u = float(r[h.index(k)]), units[h.index(k)]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
            traffic += gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    tj = os.path.join(P, f"{tag}_traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d["c3"] = int(traffic)
    json.dump(d, open(tj, "w"))
    print("GRU kernels:", len(rows) - 2, "launches, DRAM traffic per step", round(traffic / 1e9, 2), "GB")


# ---- other full captures (`ncu -i ... --page raw --csv` run on the GPU box): one block of key metrics per launch + a
# duration-weighted tensor-pipe figure -----------------------------------------------------------------------------------
def summarize_raw(raw_name, out_name, header, keyfn=lambda nm: "all"):
    raw = os.path.join(G, raw_name)
    if not os.path.exists(raw):
        return
    rows = list(csv.reader(io.StringIO(open(raw).read())))
    h, units = rows[0], rows[1]
    keep = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}
    tens, seen = {}, {}
    for r in rows[2:]:
        nm = r[h.index("Kernel Name")]
        try:
            dur = float(r[h.index("gpu__time_duration.sum")]) * tscale.get(units[h.index("gpu__time_duration.sum")], 1.0)
            tp = float(r[h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")])
            a = tens.setdefault(keyfn(nm), [0.0, 0.0, 0]); a[0] += dur * tp; a[1] += dur; a[2] += 1
        except Exception:
            pass
    with open(os.path.join(P, out_name), "w") as f:
        f.write(header)
        for kk, (num, den, n) in tens.items():
            f.write(f"# {kk}: {n} launches, {den / 1e3:.3f} ms, duration-weighted sm__pipe_tensor_cycles_active (pct of peak, elapsed) = {num / max(den, 1e-9):.1f} %\n")
        for r in rows[2:]:
            nm = r[h.index("Kernel Name")]
            key = (nm, r[h.index("launch__grid_size")] if "launch__grid_size" in h else "")
            seen[key] = seen.get(key, 0) + 1
            if seen[key] > 3:                      # at most three launches per (kernel, grid) in the tracked summary
                continue
            f.write(f"--- {nm[:140]}\n")
            for k in keep:
                if k in h: f.write(f"{k:84s} {r[h.index(k)]:>16s} {units[h.index(k)]}\n")
    print(out_name, {k: (round(v[1] / 1e3, 3), round(v[0] / max(v[1], 1e-9), 1)) for k, v in tens.items()})


summarize_raw(f"{tag}_gemm2_c3_raw.csv", f"{tag}_ncu_gemm2_c3_bf16.txt",
              "# ncu --set full --clock-control none -k regex:tc_gemm2, bench.py --workload c3: the CTA-pair GEMM launches (fn_tc_gemm2.cu) of ONE train step\n"
              "# (weight gradients dW_hh / dW_ih with split-K, the wavefront's per-segment x W^T / dy W products, logits and its gradients)\n",
              lambda nm: "tc_gemm2_kernel<176>" if "176" in nm else "tc_gemm2_kernel<256>")
summarize_raw(f"{tag}_gru_x3_c2_raw.csv", f"{tag}_ncu_gru_x3_c2.txt",
              "# ncu --set full --clock-control none -k regex:gru_tc_kernel, bench.py --workload c2_x3 (B=64,T=256,H=512, bf16x3 = fp32 parity on the tensor cores):\n"
              "# the GRU launches of ONE train step (hi/lo bf16 operand planes, three tensor-core products per algorithmic product)\n",
              lambda nm: "bwd" if __import__("re").search(r"gru_tc_kernel<\s*\d+,\s*\d+,\s*(1|true)", nm) else "fwd")
summarize_raw(f"{tag}_decode_c5_raw.csv", f"{tag}_ncu_decode_c5.txt",
              "# ncu --set full --clock-control none -k regex:decode_tc, bench.py --workload c5: ONE greedy-decode launch (256 sequences x 512 steps, H=1024, bf16)\n")
