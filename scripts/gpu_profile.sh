# ncu evidence for profiles/: (1) launch list of OUR kernels over one bench run, (2) full capture
# of the dominant kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
K='regex:carved_attn|select_blocks|pooled_scores|select_rows|block_pool|onehot_to_bits|gather_rows|hy_prologue'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 64 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:carved_attn -s 2 -c 1 \
  -o gpurun_out/attn_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
