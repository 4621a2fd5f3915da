# launch list of one train step (device time per launch, serialised / cold cache: compare SHARES) and a full capture of the GRU kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 1650 -c 560 --csv --log-file gpurun_out/r01_launches_c3_bf16.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 6 -c 6 -o gpurun_out/r01_gru_tc_c3_final python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 40 -c 12 -o gpurun_out/r01_tc_gemm_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
