mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --dit-loop 23 2>&1 | tail -1 | tee gpurun_out/bench_ditloop.log
timeout 600 python bench.py --workload wan1.3b --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_wan.log
