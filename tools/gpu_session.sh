set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest6.log 2>&1
tail -15 gpurun_out/pytest6.log
python tools/gemm_bench.py 20 2>&1 | tail -4
timeout 600 python tools/parity_report.py --config c1 c1b --pairs 2 --tag v6 > gpurun_out/parity_v6.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.log 2>&1; tail -c 1200 gpurun_out/bench_v6.log
