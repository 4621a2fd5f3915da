# Round profile of config 3 (run under gpurun; raw outputs -> gpurun_out/, summarised by tools/summarize_profiles.py):
#  1. launch list of a train step (device time per launch, serialised / cold cache: compare SHARES)
#  2. ncu --set full of the 12 GRU launches of one step (6 forward, 6 backward)
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_c3_bf16.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 12 -c 12 -f -o gpurun_out/${TAG}_gru_tc_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
