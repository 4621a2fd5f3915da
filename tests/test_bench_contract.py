"""bench.py contract on CPU: the reference arm (the oracle port timed on the host cores) prints exactly ONE JSON line on
stdout with the keys the driver reads.  The GPU arm needs a B200 and is exercised by the driver itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    # config 1 (vanilla VAE, B=4, T=128, H=256) keeps the CPU sample to about a second
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "sequences/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_workload_table_is_consistent():
    """Every train workload has a CPU-reference sample batch; the named BASELINE configs map to the modes DESIGN.md states:
    config 2 ("fp32") = the fp32 parity bar on the tensor cores (bf16x3), config 3 = bf16; both arms share one `config`."""
    sys.path.insert(0, ROOT)
    import bench
    train = [w for w in bench.WORKLOADS if not w.startswith("c5")]
    assert all(w in bench.SAMPLE_BATCH for w in train), [w for w in train if w not in bench.SAMPLE_BATCH]
    assert bench.WORKLOADS["c2"][-1] == "bf16x3" and bench.WORKLOADS["c2_f32"][-1] == "f32" and bench.WORKLOADS["c3"][-1] == "bf16"
    assert bench.WORKLOADS["c2"][1:4] == (64, 256, 512) and bench.WORKLOADS["c3"][1:4] == (256, 512, 1024)
    for w in train:
        cfg = bench.workload_config(w, 2)
        assert cfg["global_batch"] == 2 * bench.WORKLOADS[w][1] and cfg["parallelism"] == "dp2" and w in cfg["workload"]
    # algorithmic FLOP bookkeeping of the roofline: 3 x (54 H^2 + 2 H V) per token, 96 H^2 of it in the GRU kernels
    assert bench.flops_per_token(1024) == 3.0 * (54.0 * 1024 * 1024 + 2.0 * 1024 * 342)
    assert bench.gru_flops_per_token(1024) == 96.0 * 1024 * 1024
