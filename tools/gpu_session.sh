set -x
mkdir -p gpurun_out
python tools/gemm_bench.py 20 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest9.log 2>&1; tail -4 gpurun_out/pytest9.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v9.log 2>&1; tail -c 300 gpurun_out/bench_v9.log
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_v9.log 2>&1; tail -c 1500 gpurun_out/bench_c4_v9.log
