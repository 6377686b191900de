# round-end evidence: all config shapes, launch list of our kernels, DiT-loop harness
bash scripts/gpu_configs.sh
K='regex:carved_attn|select_blocks|pooled_scores|select_rows|block_pool|onehot_to_bits|gather_rows|hy_prologue'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --dit-loop 23 2>&1 | tail -1 > gpurun_out/bench_ditloop.json
python -c "import json; d=json.load(open('gpurun_out/bench_ditloop.json')); print('dit_loop', d.get('dit_loop'))"
