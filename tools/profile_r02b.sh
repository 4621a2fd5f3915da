# (ncu cannot replay cooperative cluster launches: the profile runs use FN_GRU2_COOP=0 = the same kernels and cluster
#  shape without the cooperative attribute; the grid fits the machine, so every CTA is resident either way)
TAG=r02
export FN_GRU2_COOP=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches_c3_bf16.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:gru2_ -s 36 -c 36 -f -o /tmp/${TAG}_gru_tc_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/${TAG}_gru_tc_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_gru_tc_c3_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:decode_tc -c 1 -f -o /tmp/${TAG}_decode_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
ncu -i /tmp/${TAG}_decode_c5.ncu-rep --page raw --csv > gpurun_out/${TAG}_decode_c5_raw.csv 2>/dev/null
grep -E "ERROR" gpurun_out/ncu_list.log gpurun_out/ncu_full.log gpurun_out/ncu_dec.log | head -5
ls -la gpurun_out/ /tmp/*.ncu-rep
