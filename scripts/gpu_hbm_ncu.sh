# ncu DRAM counters for the HBM-bound kernels of the path (hy_prologue, block_pool, gather_rows, selection)
mkdir -p gpurun_out
K='regex:hy_prologue|block_pool|gather_rows|pooled_scores|select_rows'
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k "$K" -c 70 --csv --log-file gpurun_out/hbm_ncu.csv python scripts/gpu_hbm_kernels.py > gpurun_out/hbm_under_ncu.log 2>&1
tail -3 gpurun_out/hbm_under_ncu.log
