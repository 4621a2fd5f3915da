timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gru_group_bf16" 2>&1 | tail -3
for r in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown 2>/dev/null | tail -1 > gpurun_out/c3_k.json
python - <<'PY'
import json,sys
d=json.load(open("gpurun_out/c3_k.json")); b=d["breakdown_ms_per_step"]
print("ms/step", d["ms_per_step"], "gru fwd", b["fn_gru_seq_fwd_bf16"][0], "bwd", b["fn_gru_seq_bwd_bf16"][0], d["last_step_outputs"][:2])
PY
done
