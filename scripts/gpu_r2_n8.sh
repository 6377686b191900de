#!/bin/bash
# 8 GPUs: BASELINE configs 2/3/5 under Ulysses-8 (fused), e2e pipelined; + the 2-GPU-needing test
W=${1:-8}
for wlname in hy720p hy_turbo_s0 hy_i2v; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $W --steps 10 --warmup 3 --workload $wlname > gpurun_out/bench_r2_n${W}_$wlname.json 2> gpurun_out/bench_r2_n${W}_$wlname.err
tail -n 1 gpurun_out/bench_r2_n${W}_$wlname.json | cut -c1-1500; tail -2 gpurun_out/bench_r2_n${W}_$wlname.err
done
JENGA_ULYSSES=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $W --steps 10 --warmup 3 > gpurun_out/bench_r2_n${W}_hy720p_nccl.json 2> gpurun_out/bench_r2_n${W}_hy720p_nccl.err
tail -n 1 gpurun_out/bench_r2_n${W}_hy720p_nccl.json | cut -c1-900
timeout 600 python bench.py --no-cpu --no-gpu-reference --dit-blocks none > gpurun_out/bench_r2_n1_same_box_as_n${W}.json 2>/dev/null; tail -n 1 gpurun_out/bench_r2_n1_same_box_as_n${W}.json | cut -c1-700
for wlname in hy_turbo_s0 hy_i2v; do
timeout 600 python bench.py --no-cpu --no-gpu-reference --dit-blocks none --workload $wlname > gpurun_out/bench_r2_n1_${wlname}_same_box.json 2>/dev/null; tail -n 1 gpurun_out/bench_r2_n1_${wlname}_same_box.json | cut -c1-500
done
