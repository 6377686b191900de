#!/bin/bash
timeout 900 python -m pytest tests/test_fp8_gpu.py tests/test_prologue_gpu.py tests/test_select_gpu.py tests/test_operator_gpu.py tests/test_dropin_blocks_gpu.py -m gpu -q -s > gpurun_out/fp8_tests.log 2>&1; echo rc=$? >> gpurun_out/fp8_tests.log
grep -n "fp8 P.V\|passed\|failed\|FAILED\|rc=\|Error" gpurun_out/fp8_tests.log | tail -20
python scripts/gpu_hbm_kernels.py 2>&1 | head -6
JENGA_PROLOGUE_SPLIT=1 python scripts/gpu_hbm_kernels.py 2>&1 | head -2
timeout 600 python bench.py --no-cpu --no-gpu-reference --no-e2e --dit-blocks none --pv-fp8 > gpurun_out/bench_fp8.json 2> gpurun_out/bench_fp8.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_fp8.json") if l.startswith("{")][-1]
    print("bf16 ms", round(d["ms_per_step"],3), "fp8 variant", d.get("fp8_pv_variant"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_fp8.err").read()[-1500:])
PY
