for g in 1 0; do
JENGA_ULYSSES_GROUPS=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_g$g.json 2> gpurun_out/bench_r2_n2_g$g.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_r2_n2_g$g.json") if l.startswith("{")][-1]
    e=d["e2e"]; print("groups=$g", "ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), e["matches_device_resident_result"], e["api"][:70])
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_r2_n2_g$g.err").read()[-1500:])
PY
done
timeout 600 python -m pytest tests/test_ulysses_gpu.py -m gpu -q 2>&1 | tail -1
