#!/bin/bash
# round 2, first GPU call: reference-backed parity, Triton ISA probe, full bench with gpu_reference
mkdir -p gpurun_out
rm -f gpurun_out/reference_parity.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python scripts/gpu_ref_probe.py > gpurun_out/ref_probe.log 2>&1
timeout 1500 python -m pytest tests/test_reference_gpu.py -q -s -m gpu > gpurun_out/ref_tests.log 2>&1
echo "ref tests rc=$?" >> gpurun_out/ref_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -c 3000 gpurun_out/ref_tests.log
cat gpurun_out/ref_probe.log | tail -30
cat gpurun_out/bench_r2a.json
