# full ncu capture of one attention-kernel generation: bash scripts/gpu_profile_gen.sh v4
mkdir -p gpurun_out
[ -n "$1" ] && export JENGA_ATTN_KERNEL=$1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:carved_attn -s 2 -c 1 \
  -o gpurun_out/attn_full_$1 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_$1.log 2>&1
ls -la gpurun_out | tail -3
