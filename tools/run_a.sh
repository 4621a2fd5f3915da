timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "latent_block or softmax or clip_adam or nll" 2>&1 | tail -4
timeout 600 python tools/bw_bench.py > gpurun_out/bw1.json 2> gpurun_out/bw1.err; tail -3 gpurun_out/bw1.err
