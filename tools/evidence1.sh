source tools/sweep_gru2.sh
run "FN_WAVEFRONT_SEGMENTS=4" seg4
run "FN_WAVEFRONT_SEGMENTS=8" seg8
run "FN_WAVEFRONT_SEGMENTS=16" seg16
mkdir -p gpurun_out/ev
tools/ubench_tmalat.bin > gpurun_out/ev/ubench_tmalat.txt 2>&1
tools/ubench_mma2.bin > gpurun_out/ev/ubench_mma2.txt 2>&1
bash tools/ubench_pair.sh > gpurun_out/ev/ubench_pair.txt 2>&1
python tools/gru2_timeline.py 256 128 1024 4 > gpurun_out/ev/timeline_fwd_c3.txt 2>&1
