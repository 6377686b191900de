#!/bin/bash
# every BASELINE.json config shape on one GPU, with the unmodified reference operator timed beside it
for wlname in wan1.3b hy720p hy_turbo_s0 wan14b hy_i2v; do
timeout 900 python bench.py --workload $wlname --no-cpu --dit-blocks none --steps 10 > gpurun_out/bench_r2_cfg_$wlname.json 2> gpurun_out/bench_r2_cfg_$wlname.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_r2_cfg_$wlname.json") if l.startswith("{")][-1]
    g=d.get("gpu_reference",{}); r=d.get("roofline",{}); e=d.get("e2e",{})
    print("$wlname", "| ours ms", round(d["ms_per_step"],3), "attn", round(r.get("ms_per_launch",0),3), "frac", round(r.get("frac",0),3), "e2e", round(e.get("ms_per_step",0),2),
          "| ref total", round(g.get("total_ms",0),2), "mask", round(g.get("mask_ms",0),2), "triton", round(g.get("triton_ms",0),2), "fa2", g.get("fa2_ms"), "tiles ref/ours", g.get("live_tiles"), d["config"]["live_tiles"], "| x", round(d.get("vs_gpu_reference") or 0,2), g.get("unavailable",""))
except Exception as ex:
    print("$wlname FAILED", ex); print(open("gpurun_out/bench_r2_cfg_$wlname.err").read()[-1000:])
PY
done
timeout 900 python bench.py --drop 0.8 --no-cpu --dit-blocks none --no-e2e --steps 10 > gpurun_out/bench_r2_cfg_hy720p_drop08.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_r2_cfg_hy720p_drop08.json") if l.startswith("{")][-1]
g=d.get("gpu_reference",{})
print("hy720p drop0.8 | ours ms", round(d["ms_per_step"],3), "| ref", round(g.get("total_ms",0),2), "tiles", g.get("live_tiles"), d["config"]["live_tiles"], "x", round(d.get("vs_gpu_reference") or 0,2))
PY
