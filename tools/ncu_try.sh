for cfg in "FN_GRU2_MC=1 FN_GRU2_MC_BWD=1" "FN_GRU2_COOP=0" "FN_GRU2_COOP=0 FN_GRU2_MC=1 FN_GRU2_MC_BWD=1" "FN_GRU2_MC=2"; do
  env $cfg ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gru2_ -c 4 --csv --log-file gpurun_out/ncu_try.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_try.log 2>&1
  echo "== $cfg: $(grep -c gru2_ gpurun_out/ncu_try.csv) rows; $(grep -E 'ERROR|LaunchFailed' gpurun_out/ncu_try.log | head -2 | tr '\n' ' ')"
done
