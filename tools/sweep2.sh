source tools/sweep_gru2.sh
run "FN_GRU_V2=1" d2b_k4s2w2
run "FN_GRU2_KCH=4 FN_GRU2_S=2 FN_GRU2_WST=3" d2b_k4s2w3
run "FN_GRU2_KCH=2 FN_GRU2_S=4 FN_GRU2_WST=4" d2b_k2s4w4
run "FN_GRU2_KCH=2 FN_GRU2_S=3 FN_GRU2_WST=2" d2b_k2s3w2
run "FN_GRU2_KCH=2 FN_GRU2_S=5 FN_GRU2_WST=5" d2b_k2s5w5
grep -h plan2 gpurun_out/sw_d2b*.err | sort | uniq
