mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for g in 6 8 12 24; do
  JENGA_E2E_GROUPS=$g timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups $g e2e ms',round(d['e2e']['ms_per_step'],2),'ok',d['e2e']['matches_device_resident_result'],'dev ms',round(d['ms_per_step'],2))" | tee -a gpurun_out/sweep.log
done
