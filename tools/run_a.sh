timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in c2_bf16 c5 c2; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/b_$w.json
python - $w <<'PY'
import json,sys
d=json.load(open("gpurun_out/b_%s.json" % sys.argv[1]))
print(sys.argv[1], d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
