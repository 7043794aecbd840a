set -x
timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -3
NMRF_B200_LIB=nmrf_b200/libnmrf_b200_e1.so timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -3
NMRF_B200_LIB=nmrf_b200/libnmrf_b200_e2.so timeout 300 python tools/gemm_bench.py 20 2>&1 | tail -3
