mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for v in "" qkfull; do
  if [ -n "$v" ]; then export JENGA_B200_LIB=$PWD/jenga_b200/_C/libjenga_b200.$v.so; else unset JENGA_B200_LIB; fi
  echo "== variant ${v:-default}" | tee -a gpurun_out/sweep.log
  timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_operator_gpu.py -q -x 2>&1 | tail -2 | tee -a gpurun_out/sweep.log
  for i in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/sweep.log
  done
done
