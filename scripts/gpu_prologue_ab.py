"""A/B of the two forms of the HunyuanVideo prologue at HY-720p (classic register-staged kernel vs the
bulk-async persistent kernel, the default; JENGA_PROLOGUE=classic selects the other): CUDA-event time, GB/s against MEASURED_PEAKS.json."""
import json, os, sys, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from jenga_b200.hyvideo import attention_prologue
dev = torch.device("cuda", 0)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
L, T, H = 115200, 256, 24
S = L + T
img = torch.randn(1, L, 3 * H * 128, device=dev).bfloat16()
txt = torch.randn(1, T, 3 * H * 128, device=dev).bfloat16()
ws = [(1 + 0.1 * torch.randn(128, device=dev)).bfloat16() for _ in range(4)]
cos = torch.rand(L, 128, device=dev); sin = torch.rand(L, 128, device=dev)
index = torch.randperm(L, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
byt = 2 * (3 * S * H * 128 * 2) + 2 * L * 128 * 4
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n
for label, idx in (("identity rope rows", None), ("rope_index = random permutation", index)):
    outs = {}
    for mode in ("classic", "bulk"):
        os.environ["JENGA_PROLOGUE"] = mode
        fn = lambda: attention_prologue(img, txt, H, *ws, eps=1e-6, freqs_cis=(cos, sin), rope_index=idx)
        ms = timeit(fn)
        outs[mode] = fn()
        print(f"{label:34s} {mode:8s} {ms:7.3f} ms  {byt/ms/1e6:8.1f} GB/s  {byt/ms/1e6/peak:5.2f} of measured HBM peak", flush=True)
    a, b = outs["classic"], outs["bulk"]
    same = all(torch.equal(x.view(torch.int16), y.view(torch.int16)) for x, y in zip(a[:3], b[:3]))
    pool_same = (a[3][0].view(torch.int16) == b[3][0].view(torch.int16)).float().mean().item()
    print(f"   q,k,v bit-identical: {same}; pooled q identical fraction {pool_same:.5f}", flush=True)
