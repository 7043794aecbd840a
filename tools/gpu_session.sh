set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest4.log 2>&1
tail -25 gpurun_out/pytest4.log
timeout 300 python tools/precision_probe.py v4 > gpurun_out/precision_v4.log 2>&1; head -12 gpurun_out/precision_v4.log
timeout 600 python tools/parity_report.py --config c1 c1b --pairs 3 --tag v4 > gpurun_out/parity_v4.log 2>&1
timeout 600 python tools/parity_report.py --config c2 --pairs 1 --tag v4 > gpurun_out/parity_c2_v4.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4.log 2>&1; tail -c 1500 gpurun_out/bench_v4.log
