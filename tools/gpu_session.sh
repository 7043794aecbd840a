set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "conv2d" > gpurun_out/s35_pytest.log 2>&1; tail -15 gpurun_out/s35_pytest.log
