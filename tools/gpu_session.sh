set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config c4 --steps 8 --warmup 3 > $O/bench_c4_8gpu.log 2>&1; tail -c 500 $O/bench_c4_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_c1_8gpu.log 2>&1; tail -c 400 $O/bench_c1_8gpu.log
