#!/bin/bash
# 2 GPUs: whole GPU suite (incl. the 2-GPU Ulysses test) + N=2 bench fused / nccl + N=1 on the same box
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2d.log 2>&1; echo rc=$? >> gpurun_out/gpu_tests_r2d.log
tail -5 gpurun_out/gpu_tests_r2d.log
for mode in fused nccl; do
JENGA_ULYSSES=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_$mode.json 2> gpurun_out/bench_r2_n2_$mode.err
tail -n 1 gpurun_out/bench_r2_n2_$mode.json | cut -c1-1800; tail -3 gpurun_out/bench_r2_n2_$mode.err
done
timeout 600 python bench.py --no-cpu --dit-blocks 4,8 > gpurun_out/bench_r2_n1_same_box.json 2> gpurun_out/bench_r2_n1_same_box.err; tail -n 1 gpurun_out/bench_r2_n1_same_box.json | cut -c1-3000
