source tools/sweep_gru2.sh
run "FN_GRU_V2=0" v1
run "FN_GRU_V2=1" k2s3w2
run "FN_GRU2_KCH=1 FN_GRU2_S=5 FN_GRU2_WST=3" k1s5w3
run "FN_GRU2_KCH=1 FN_GRU2_S=6 FN_GRU2_WST=4" k1s6w4
run "FN_GRU2_KCH=2 FN_GRU2_S=2 FN_GRU2_WST=2" k2s2w2
run "FN_GRU2_KCH=2 FN_GRU2_S=4 FN_GRU2_WST=2" k2s4w2
run "FN_GRU2_KCH=4 FN_GRU2_S=2 FN_GRU2_WST=2" k4s2w2
grep plan2 gpurun_out/sw_*.err | sort | uniq | head -20
