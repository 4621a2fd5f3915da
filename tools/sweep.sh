for cs in 1 2 4; do
  FN_GRU_CLUSTER=$cs timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --breakdown 2>&1 | tail -1 > gpurun_out/c3_cs_$cs.json
done
FN_GRU_RING_KB=112 FN_GRU_WRING_KB=24 timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --breakdown 2>&1 | tail -1 > gpurun_out/c3_cs_r112.json
timeout 300 python bench.py --workload c2_bf16 --steps 5 --warmup 3 --no-cpu-baseline --breakdown 2>&1 | tail -1 > gpurun_out/c2bf16_now.json
