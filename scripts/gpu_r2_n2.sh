#!/bin/bash
# 2 GPUs: Ulysses tests (head-group pipelining, double-buffered output) + N=2 bench fused / nccl
timeout 900 python -m pytest tests/test_ulysses_gpu.py -m gpu -q -x > gpurun_out/ulysses_n2_tests.log 2>&1; echo rc=$? >> gpurun_out/ulysses_n2_tests.log
tail -5 gpurun_out/ulysses_n2_tests.log
for mode in fused nccl; do
JENGA_ULYSSES=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_$mode.json 2> gpurun_out/bench_r2_n2_$mode.err
tail -c 1800 gpurun_out/bench_r2_n2_$mode.json; tail -3 gpurun_out/bench_r2_n2_$mode.err
done
timeout 600 python bench.py --no-cpu --no-gpu-reference --dit-blocks none > gpurun_out/bench_r2_n1_same_box.json 2>&1; tail -c 600 gpurun_out/bench_r2_n1_same_box.json
