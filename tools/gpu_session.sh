set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "token_gemm or row_stats or mlp_chain" > gpurun_out/pytest10a.log 2>&1; tail -12 gpurun_out/pytest10a.log
timeout 120 python tools/gemm_bench.py 20 2>&1 | tail -4
NMRF_B200_GEMM_RA=0 timeout 120 python tools/gemm_bench.py 20 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest10.log 2>&1; tail -6 gpurun_out/pytest10.log
timeout 300 python tools/precision_probe.py v10 > gpurun_out/precision_v10.log 2>&1; head -12 gpurun_out/precision_v10.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10.log 2>&1; tail -c 300 gpurun_out/bench_v10.log
