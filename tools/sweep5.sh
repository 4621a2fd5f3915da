source tools/sweep_gru2.sh
run "FN_GRU_V2=3" bwd_mc4
run "FN_GRU_V2=3 FN_GRU2_MC_BWD=2" bwd_mc2
run "FN_GRU_V2=3 FN_GRU2_MC_BWD=1" bwd_mc1
grep -h plan2 gpurun_out/sw_bwd*.err | sort | uniq
