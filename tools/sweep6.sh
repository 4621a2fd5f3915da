source tools/sweep_gru2.sh
run "FN_GRU2_S_BWD=3" b_k4s3w2
run "FN_GRU2_KCH_BWD=2 FN_GRU2_S_BWD=6 FN_GRU2_WST_BWD=4" b_k2s6w4
run "FN_GRU2_KCH_BWD=2 FN_GRU2_S_BWD=5 FN_GRU2_WST_BWD=3" b_k2s5w3
run "FN_GRU2_KCH_BWD=2 FN_GRU2_S_BWD=4 FN_GRU2_WST_BWD=2" b_k2s4w2
run "FN_GRU2_S_BWD=3 FN_GRU2_MC_BWD=2" b_k4s3w2_mc2
grep -h "plan2 bwd" gpurun_out/sw_b_*.err | sort | uniq
