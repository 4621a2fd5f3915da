source tools/sweep_gru2.sh
run "FN_GRU2_GEO=64" p_g64
run "FN_GRU2_KCH=2 FN_GRU2_S=5 FN_GRU2_WST=3" p_k2s5w3
run "FN_GRU2_MC=2" p_mc2
run "FN_GRU2_MC=1" p_mc1
run "FN_GRU2_MC_BWD=4" p_bmc4
run "FN_GRU2_MC_BWD=2" p_bmc2
run "FN_GRU2_KCH_BWD=2 FN_GRU2_S_BWD=6 FN_GRU2_WST_BWD=4" p_bk2s6
