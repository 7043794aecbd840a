set -x
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -4
NMRF_B200_LIB=nmrf_b200/libnmrf_b200_nohint.so timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s32_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s32_bench.log
NMRF_B200_LIB=nmrf_b200/libnmrf_b200_nohint.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s32_bench_nohint.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s32_bench_nohint.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s32_pytest.log 2>&1; tail -3 gpurun_out/s32_pytest.log
