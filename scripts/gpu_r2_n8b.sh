#!/bin/bash
# 8 GPUs, second pass: sub-group rule (>= 12 waves per launch) -> 1 group per rank at N=8
W=${1:-8}
for wlname in hy720p hy_turbo_s0 hy_i2v; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $W --steps 10 --warmup 3 --workload $wlname > gpurun_out/bench_r2_n${W}b_$wlname.json 2> gpurun_out/bench_r2_n${W}b_$wlname.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_r2_n${W}b_$wlname.json") if l.startswith("{")][-1]
    e=d.get("e2e") or {}
    print("$wlname N=$W", "ms", round(d["ms_per_step"],3), "TF/s", round(d["value"],1), "| e2e ms", round(e.get("ms_per_step",0),3), e.get("matches_device_resident_result"), e.get("api","")[:60])
except Exception as ex:
    print("$wlname FAILED", ex); print(open("gpurun_out/bench_r2_n${W}b_$wlname.err").read()[-1200:])
PY
done
timeout 600 python bench.py --no-cpu --no-gpu-reference --dit-blocks none > gpurun_out/bench_r2_n1_same_box_as_n${W}b.json 2>/dev/null; tail -n 1 gpurun_out/bench_r2_n1_same_box_as_n${W}b.json | cut -c1-400
