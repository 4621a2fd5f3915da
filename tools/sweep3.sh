source tools/sweep_gru2.sh
run "FN_GRU2_GEO=64" g64_k2s3w2
run "FN_GRU2_GEO=64 FN_GRU2_KCH=2 FN_GRU2_S=2 FN_GRU2_WST=2" g64_k2s2w2
run "FN_GRU2_GEO=64 FN_GRU2_KCH=2 FN_GRU2_S=4 FN_GRU2_WST=2" g64_k2s4w2
run "FN_GRU2_GEO=64 FN_GRU2_KCH=4 FN_GRU2_S=2 FN_GRU2_WST=1" g64_k4s2w1
run "FN_GRU2_GEO=32" g32_k4s2w2
grep -h plan2 gpurun_out/sw_g*.err | sort | uniq
