mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ulysses_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/gpu_tests_n2.log
for m in fused nccl; do
JENGA_ULYSSES=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', 'value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['config']['parallelism'])" | tee -a gpurun_out/bench_n2.log
done
