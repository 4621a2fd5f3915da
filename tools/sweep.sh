for cfg in "48 24" "64 32" "96 32" "128 24" "64 48"; do
  set -- $cfg
  FN_GRU_RING_KB=$1 FN_GRU_WRING_KB=$2 timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/c5_sw_$1_$2.json
done
