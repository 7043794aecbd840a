set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv2d or encoder or instnorm" > gpurun_out/s46_pytest_a.log 2>&1; tail -8 gpurun_out/s46_pytest_a.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s46_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s46_bench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s46_pytest.log 2>&1; tail -3 gpurun_out/s46_pytest.log
