mkdir -p gpurun_out
export JENGA_ATTN_KERNEL=v3
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:carved_attn -s 2 -c 1 \
  -o gpurun_out/attn_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out | tail -3
