#!/bin/bash
# usage: tools/sweep_gru2.sh "<env assignments>" tag  -> gpurun_out/sw_<tag>.json (c3 bench with per-call breakdown)
run() { env $1 FN_GRU_VERBOSE=1 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline --no-gpu-reference > gpurun_out/sw_$2.json 2> gpurun_out/sw_$2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_$2.json").read().strip().splitlines()[-1])
    b=d["breakdown_ms_per_step"]
    print("$2", d["ms_per_step"], "gru", d["roofline"]["ms_per_step_in_kernel"], "fwd", b.get("fn_gru_seq_fwd_bf16"), "bwd", b.get("fn_gru_seq_bwd_bf16"))
except Exception as e:
    print("$2 FAILED", e); print(open("gpurun_out/sw_$2.err").read()[-800:])
PY
}
