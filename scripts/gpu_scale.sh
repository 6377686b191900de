mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n$N.log
