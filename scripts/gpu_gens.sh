# parity tests + bench for several attention-kernel generations: bash scripts/gpu_gens.sh v2 v4 v5
mkdir -p gpurun_out; rm -f gpurun_out/gens.log
for g in "$@"; do
  if [ "$g" = "v2" ]; then unset JENGA_ATTN_KERNEL; else export JENGA_ATTN_KERNEL=$g; fi
  timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_fullsize_gpu.py tests/test_operator_gpu.py -q -x 2>&1 | tail -2 | tee -a gpurun_out/gens.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$g value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/gens.log
done
