mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --dit-loop 2 2>&1 | tail -1 | tee gpurun_out/bench_full.log
timeout 300 python scripts/gpu_hbm_kernels.py 2>&1 | tail -4 | tee gpurun_out/hbm_kernels.log
K='regex:carved_attn|select_blocks|block_pool|onehot_to_bits|gather_rows|hy_prologue'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 64 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
