#!/bin/bash
# ncu launch list (per-launch durations) of one default bench run: the dominant kernel's SHARE of the step
mkdir -p gpurun_out
K='regex:carved_attn|select_rows|pooled_scores|block_pool|onehot_to_bits|gather_rows|hy_prologue|ln_modulate|gate_residual|gelu_tanh'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gpu-reference --dit-blocks none > gpurun_out/bench_under_ncu_r2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].split("::")[-1]
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u.startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':40s} launches   total ms   share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {n:8d} {ms:10.3f} {100*ms/tot:7.2f}%")
PY
