source tools/sweep_gru2.sh
run "FN_GRU2_GEO=32 FN_GRU2_MC=1" mc1_g32
run "FN_GRU2_GEO=32 FN_GRU2_MC=2" mc2_g32
run "FN_GRU2_GEO=32 FN_GRU2_MC=4" mc4_g32
run "FN_GRU2_GEO=64 FN_GRU2_MC=2" mc2_g64
run "FN_GRU2_GEO=64 FN_GRU2_MC=2 FN_GRU2_KCH=4 FN_GRU2_S=2 FN_GRU2_WST=2" mc2_g64_k4
grep -h plan2 gpurun_out/sw_mc*.err | sort | uniq
