set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_c1.log 2>&1; tail -c 300 $O/bench_c1.log
for c in c1b c2 c4; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.log 2>&1; tail -c 150 $O/bench_$c.log; done
timeout 1200 python tools/parity_report.py --config c1 c1b c2 --tag final > $O/parity_final.log 2>&1; tail -3 $O/parity_final.log
cp gpurun_out/parity_*_final.json $O/ 2>/dev/null
timeout 300 python tools/conv_bench.py 20 > $O/conv_bench.log 2>&1; cat $O/conv_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_forward.csv python tools/ncu_target.py 1 > $O/ncu_list.log 2>&1; tail -1 $O/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:token_gemm_tc6 -c 8 -f -o $O/prof_conv python tools/ncu_target.py 1 > $O/ncu_conv.log 2>&1; tail -1 $O/ncu_conv.log
