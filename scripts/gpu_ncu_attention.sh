#!/bin/bash
# ncu --set full of generation 7 (and generation 2 as control) on the headline workload; read here with --page source
mkdir -p gpurun_out
for gen in v7 v2; do
JENGA_ATTN_KERNEL=$gen timeout 1200 ncu --set full --clock-control none --import-source on -k regex:carved_attn -s 2 -c 1 \
  -o gpurun_out/r02_attn_$gen -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gpu-reference --dit-blocks none > gpurun_out/ncu_$gen.log 2>&1
tail -2 gpurun_out/ncu_$gen.log
done
ls -la gpurun_out/*.ncu-rep
