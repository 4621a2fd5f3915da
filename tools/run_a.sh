bash tools/profile.sh r01
python tools/bw_bench.py > gpurun_out/r01_bandwidth.json 2>/dev/null
python bench.py > gpurun_out/r01_bench_c3.json 2> gpurun_out/r01_bench_c3.err; tail -c 600 gpurun_out/r01_bench_c3.json
