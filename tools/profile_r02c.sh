# Final round-2 profile (run under gpurun; raw outputs -> gpurun_out/, summarised by tools/summarize_profiles.py r02):
#  1. launch list of a c3 train step (device time per launch, serialised / cold cache: compare SHARES)
#  2. ncu --set full: the 36 GRU launches of one c3 step; the CTA-pair GEMM (tc_gemm2) launches of that step; the bf16x3 GRU
#     launches of one c2_x3 step; the greedy-decode launch of c5
#  3. compute-sanitizer memcheck + racecheck of the new kernels (bf16x3 GRU / GEMM planes, CTA-pair GEMM) at small shapes
# (ncu cannot replay cooperative cluster launches: FN_GRU2_COOP=0 = the same kernels and cluster shape without the attribute)
TAG=r02
export FN_GRU2_COOP=0
B="--no-cpu-baseline --no-gpu-reference --no-parity-mode"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches_c3_bf16.csv python bench.py --workload c3 --steps 1 --warmup 1 $B > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:gru2_ -s 36 -c 36 -f -o /tmp/${TAG}_gru_tc_c3 python bench.py --workload c3 --steps 1 --warmup 1 $B > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/${TAG}_gru_tc_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_gru_tc_c3_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:tc_gemm2 -s 60 -c 60 -f -o /tmp/${TAG}_gemm2_c3 python bench.py --workload c3 --steps 1 --warmup 1 $B > gpurun_out/ncu_gemm2.log 2>&1
ncu -i /tmp/${TAG}_gemm2_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm2_c3_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:gru_tc_kernel -s 20 -c 20 -f -o /tmp/${TAG}_gru_x3_c2 python bench.py --workload c2_x3 --steps 1 --warmup 1 $B > gpurun_out/ncu_x3.log 2>&1
ncu -i /tmp/${TAG}_gru_x3_c2.ncu-rep --page raw --csv > gpurun_out/${TAG}_gru_x3_c2_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:decode_tc -c 1 -f -o /tmp/${TAG}_decode_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
ncu -i /tmp/${TAG}_decode_c5.ncu-rep --page raw --csv > gpurun_out/${TAG}_decode_c5_raw.csv 2>/dev/null
unset FN_GRU2_COOP
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_x3.py tests/test_gpu_ops.py -x -q -k "(x3 and (70 or 130-4)) or (tc_gemm_x3 and 300) or (tc_gemm_bf16 and (300-640 or 700-342)) or splitk" > gpurun_out/${TAG}_sanitizer_memcheck_x3_gemm2.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_x3.py tests/test_gpu_ops.py -x -q -k "(gru_group_x3 and 70) or (tc_gemm_bf16 and 300-640)" > gpurun_out/${TAG}_sanitizer_racecheck_x3_gemm2.txt 2>&1
grep -E "ERROR|==ERROR" gpurun_out/ncu_list.log gpurun_out/ncu_full.log gpurun_out/ncu_gemm2.log gpurun_out/ncu_x3.log gpurun_out/ncu_dec.log | head -5
tail -n 3 gpurun_out/${TAG}_sanitizer_memcheck_x3_gemm2.txt gpurun_out/${TAG}_sanitizer_racecheck_x3_gemm2.txt
ls -la gpurun_out/ | tail -n 20
