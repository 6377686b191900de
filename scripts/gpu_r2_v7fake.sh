#!/bin/bash
run() { # name lib gen
python - <<PY
import json, os, subprocess, sys
env = dict(os.environ, JENGA_ATTN_KERNEL="$3")
if "$2": env["JENGA_B200_LIB"] = "$2"
r = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-gpu-reference", "--no-e2e", "--dit-blocks", "none", "--steps", "8"], env=env, capture_output=True, text=True)
try:
    d=[json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")][-1]
    print("$1", "attn ms", round(d["roofline"]["ms_per_launch"],3), "TF/s", round(d["roofline"]["achieved"],1), "clock", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"])
except Exception as e:
    print("$1 FAILED", e, r.stderr[-800:])
PY
}
run gen2 "" v2
run gen7 "" v7
run gen7_burst jenga_b200/_C/libjenga_b200.burst.so v7
run gen7_qqp jenga_b200/_C/libjenga_b200.qqp.so v7
run gen2_again "" v2
