mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for v in "" g4b1 g8b1 g2b2; do
  if [ -n "$v" ]; then export JENGA_B200_LIB=$PWD/jenga_b200/_C/libjenga_b200.$v.so; else unset JENGA_B200_LIB; fi
  echo "== variant ${v:-default(g4b2)}" | tee -a gpurun_out/sweep.log
  timeout 300 python scripts/gpu_hbm_kernels.py 2>&1 | grep hy_prologue | tee -a gpurun_out/sweep.log
done
unset JENGA_B200_LIB
timeout 300 python -m pytest tests/test_prologue_gpu.py tests/test_shims_gpu.py -q 2>&1 | tail -2 | tee -a gpurun_out/sweep.log
