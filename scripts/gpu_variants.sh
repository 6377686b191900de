# bench tuning variants (libjenga_b200.<name>.so) x kernel generations: bash scripts/gpu_variants.sh "default p0 roll" "v2 v6"
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
for v in $1; do
  if [ "$v" = "default" ]; then unset JENGA_B200_LIB; else export JENGA_B200_LIB=$PWD/jenga_b200/_C/libjenga_b200.$v.so; fi
  for g in $2; do
    if [ "$g" = "v2" ]; then unset JENGA_ATTN_KERNEL; else export JENGA_ATTN_KERNEL=$g; fi
    if [ -n "$3" ]; then timeout 300 python -m pytest tests/test_attn_gpu.py -q -x 2>&1 | tail -1 | tee -a gpurun_out/variants.log; fi
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v $g attn ms',round(d['roofline']['ms_per_launch'],2),'TF/s',round(d['roofline']['achieved'],1),'clk',d['clocks']['sm_mhz'],'W',d['clocks'].get('power_w_max'))" | tee -a gpurun_out/variants.log
  done
done
