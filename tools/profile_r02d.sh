export FN_GRU2_COOP=0
B="--no-cpu-baseline --no-gpu-reference --no-parity-mode"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c3_bf16.csv python bench.py --workload c3 --steps 1 --warmup 1 $B > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:gru2_ -s 36 -c 36 -f -o /tmp/r02_gru_tc_c3 python bench.py --workload c3 --steps 1 --warmup 1 $B > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r02_gru_tc_c3.ncu-rep --page raw --csv > gpurun_out/r02_gru_tc_c3_raw.csv 2>/dev/null
ls -la gpurun_out/r02_gru_tc_c3_raw.csv gpurun_out/r02_launches_c3_bf16.csv
