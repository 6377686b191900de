#!/bin/bash
# generation 7 first light: parity + bit-equality vs generation 2, then A/B on the headline workload
timeout 1200 python -m pytest tests/test_attn_gpu.py -m gpu -q -x -k "generation7 or matches_oracle" > gpurun_out/v7_tests.log 2>&1; echo rc=$? >> gpurun_out/v7_tests.log
tail -25 gpurun_out/v7_tests.log
for gen in v2 v7 v2 v7; do
JENGA_ATTN_KERNEL=$gen timeout 600 python bench.py --no-cpu --no-gpu-reference --no-e2e --dit-blocks none --steps 10 > gpurun_out/bench_v7ab_$gen.json 2> gpurun_out/bench_v7ab_$gen.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_v7ab_$gen.json") if l.startswith("{")][-1]
    print("$gen", "step ms", round(d["ms_per_step"],3), "attn ms", round(d["roofline"]["ms_per_launch"],3), "TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],4), d["clocks"])
except Exception as e:
    print("$gen", "FAILED", e); print(open("gpurun_out/bench_v7ab_$gen.err").read()[-1500:])
PY
done
