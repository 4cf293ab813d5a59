for p in 0 1; do echo "=== PAIR=$p"; CRA5_GEMM_PAIR=$p python tools/perf_kernels.py 2>&1 | grep -v cublas | head -20; done
python tools/perf_kernels.py 2>&1 | grep -A1 "gemm" | grep cublas
