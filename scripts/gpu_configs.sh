mkdir -p gpurun_out; rm -f gpurun_out/configs.log
for w in hy_turbo_s0 hy_i2v wan14b wan1.3b; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', 'op ms',round(d['ms_per_step'],3),'TF/s',round(d['value'],1),'| attn ms',round(r['ms_per_launch'],3),'TF/s',round(r['achieved'],1),'frac',round(r['frac'],3),'| e2e ms',round(d['e2e']['ms_per_step'],2),d['e2e']['matches_device_resident_result'],'| tiles',r['live_tiles'],'TFLOP',round(d['config']['algorithmic_tflop_per_step'],2),'clk',d['clocks']['sm_mhz'])" | tee -a gpurun_out/configs.log
done
