for l in 1 2 4; do echo "lanes=$l"; FN_GEMM_LANES=$l timeout 300 python tools/gemm_bench.py 2>&1 | tail -5; done
