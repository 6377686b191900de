mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for v in "" g4b1 g2b2 g1b2; do
  if [ -n "$v" ]; then export JENGA_B200_LIB=$PWD/jenga_b200/_C/libjenga_b200.$v.so; else unset JENGA_B200_LIB; fi
  echo "== variant ${v:-default(g2b1)}" | tee -a gpurun_out/sweep.log
  timeout 300 python scripts/gpu_hbm_kernels.py 2>&1 | grep -E "hy_prologue|select" | tee -a gpurun_out/sweep.log
done
unset JENGA_B200_LIB
timeout 300 python -m pytest tests/test_prologue_gpu.py tests/test_shims_gpu.py tests/test_select_gpu.py -q 2>&1 | tail -2 | tee -a gpurun_out/sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:select_blocks" -s 3 -c 1 -o gpurun_out/sel_full -f python scripts/gpu_hbm_kernels.py > gpurun_out/aux_ncu.log 2>&1
