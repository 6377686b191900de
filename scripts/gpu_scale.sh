mkdir -p gpurun_out
N=${1:-8}; M=${2:-fused}
JENGA_ULYSSES=$M timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_$M.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N $M', 'value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['config']['parallelism'], d['clocks'])"
