# What a round-end check runs on the GPU box (gpurun -- 'bash tools/gpu_session.sh'); outputs under gpurun_out/final/
set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_c1.log 2>&1; tail -c 300 $O/bench_c1.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c1_reference.log 2>&1; tail -c 200 $O/bench_c1_reference.log
for c in c1b c2 c4; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.log 2>&1; tail -c 150 $O/bench_$c.log; done
