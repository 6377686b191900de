mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ulysses_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/gpu_tests_n2.log
