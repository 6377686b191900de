mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_v2.log
