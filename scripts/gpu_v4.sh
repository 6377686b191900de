# generation-4 attention kernel: parity tests + bench (v4 then v2 for reference)
mkdir -p gpurun_out; rm -f gpurun_out/v4.log
export JENGA_ATTN_KERNEL=${GEN:-v4}
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_fullsize_gpu.py tests/test_operator_gpu.py -q -x 2>&1 | tail -8 | tee -a gpurun_out/v4.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gen value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/v4.log
unset JENGA_ATTN_KERNEL
if [ "$1" = "both" ]; then
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2 value',round(d['value'],1),'attn TF/s',round(d['roofline']['achieved'],1),'ms',round(d['roofline']['ms_per_launch'],2),'clk',d['clocks'])" | tee -a gpurun_out/v4.log
fi
