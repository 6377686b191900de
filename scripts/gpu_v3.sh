mkdir -p gpurun_out; rm -f gpurun_out/v3.log
for g in v3; do
export JENGA_ATTN_KERNEL=$g
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_fullsize_gpu.py -q -x 2>&1 | tail -3 | tee -a gpurun_out/v3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$g value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/v3.log
done
unset JENGA_ATTN_KERNEL
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2 value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/v3.log
