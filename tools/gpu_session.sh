set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --config c3 --steps 8 --warmup 3 > $O/bench_c3_4gpu.log 2>&1; tail -c 600 $O/bench_c3_4gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_c1_4gpu.log 2>&1; tail -c 400 $O/bench_c1_4gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c1_2gpu.log 2>&1; tail -c 400 $O/bench_c1_2gpu.log
timeout 900 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_c3_1gpu.log 2>&1; tail -c 300 $O/bench_c3_1gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $O/bench_ref_2gpu.log 2>&1; tail -c 300 $O/bench_ref_2gpu.log
