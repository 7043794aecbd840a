set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "token_gemm or row_stats or encoder" > gpurun_out/s43_pytest_a.log 2>&1; tail -4 gpurun_out/s43_pytest_a.log
timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -3
NMRF_B200_GEMM_TMA=0 timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s43_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s43_bench.log; grep -o '"nmrf_token_gemm": {[^}]*}' gpurun_out/s43_bench.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s43_pytest.log 2>&1; tail -3 gpurun_out/s43_pytest.log
