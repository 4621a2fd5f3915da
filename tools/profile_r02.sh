# Round-2 profile (run under gpurun; raw outputs -> gpurun_out/, summarised by tools/summarize_profiles.py r02):
#  1. launch list of a c3 train step (device time per launch, serialised / cold cache: compare SHARES)
#  2. ncu --set full of the GRU launches of one step (CTA-pair kernels, 16-segment wavefront: 18 forward + 18 backward)
#  3. DRAM bytes + duration of the bandwidth-bound kernels (tools/bw_bench.py): vocab_*, gm_kl_*, qy_*, reparam, grad_norm, clip_adam
#  4. compute-sanitizer memcheck + racecheck of the GRU / decode kernels at a small shape
TAG=r02
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches_c3_bf16.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru2_ -s 36 -c 36 -f -o gpurun_out/${TAG}_gru_tc_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'vocab_|gm_kl|qy_|reparam|std_kl|sq_stage|reduce_stage|clip_adam|latent_reg' -c 400 --csv --log-file gpurun_out/${TAG}_bw_kernels.csv python tools/bw_bench.py > gpurun_out/${TAG}_bw_under_ncu.json 2> gpurun_out/ncu_bw.log
python tools/bw_bench.py > gpurun_out/${TAG}_bandwidth.json 2> gpurun_out/bw.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_gru_pair.py tests/test_gpu_parity_bf16.py -x -q -k "129 or 200-33 or (argmax and 128)" > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_gru_pair.py -x -q -k "200-33" > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log gpurun_out/${TAG}_sanitizer_memcheck.txt gpurun_out/${TAG}_sanitizer_racecheck.txt
