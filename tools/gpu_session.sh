set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "token_gemm or row_stats or mlp_chain or cost_volume" > gpurun_out/pytest11a.log 2>&1; tail -5 gpurun_out/pytest11a.log
timeout 120 python tools/gemm_bench.py 20 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest11.log 2>&1; tail -4 gpurun_out/pytest11.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v11.log 2>&1; tail -c 300 gpurun_out/bench_v11.log
