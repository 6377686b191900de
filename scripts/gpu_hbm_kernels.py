"""Times the HBM-bound kernels of the path at HY-720p size and prints achieved GB/s vs the
measured HBM peak (MEASURED_PEAKS.json hbm_gbs)."""
import json, sys, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from jenga_b200 import gilbert
from jenga_b200.attention import block_pool, select_blocks, neighbour_bits
from jenga_b200.hyvideo import attention_prologue, gather_tokens
dev = torch.device("cuda", 0)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
L, T, H = 115200, 256, 24
S = L + T
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = {}
img = torch.randn(1, L, 3 * H * 128, device=dev).bfloat16()
txt = torch.randn(1, T, 3 * H * 128, device=dev).bfloat16()
ws = [torch.ones(128, dtype=torch.bfloat16, device=dev) for _ in range(4)]
cos = torch.rand(L, 128, device=dev); sin = torch.rand(L, 128, device=dev)
ms = timeit(lambda: attention_prologue(img, txt, H, *ws, eps=1e-6, freqs_cis=(cos, sin)))
byt = 2 * (3 * S * H * 128 * 2) + 2 * L * 128 * 4
res["hy_prologue"] = (ms, byt)
q, k, v, pools = attention_prologue(img, txt, H, *ws, eps=1e-6, freqs_cis=(cos, sin))
ms = timeit(lambda: block_pool(k, 902)); res["block_pool(k)"] = (ms, S * H * 128 * 2)
x = torch.randn(1, L, 3072, device=dev).bfloat16()
l2h, h2l = gilbert.mapping_tensors(32, 45, 80); h2l = h2l.to(dev)
ms = timeit(lambda: gather_tokens(x, h2l)); res["gather_rows"] = (ms, 2 * L * 3072 * 2)
nbr = neighbour_bits(gilbert.block_neighbor_mapping(32, 45, 80), dev)
qp, kp = pools
ms = timeit(lambda: select_blocks(qp[:, :, :900].contiguous(), kp, n_img=900, nb=902, top_k=270, p_threshold=0.3, text_blocks=2, nbr_bits=nbr))
res["select_blocks"] = (ms, 0)
for k_, (ms, b) in res.items():
    print(f"{k_:16s} {ms:8.3f} ms  {b/ms/1e6:9.1f} GB/s  frac_of_hbm_peak {b/ms/1e6/peak:5.2f}" if b else f"{k_:16s} {ms:8.3f} ms")

# ---- round 2: f-1 chains, residual cache, device gilbert (algorithmic bytes = one read + one write of the tensors)
from jenga_b200 import dense as DN, stepcache as SC
xs = torch.randn(1, S, 3072, device=dev).bfloat16()
sh = torch.randn(1, 3072, device=dev).bfloat16(); sc = torch.randn(1, 3072, device=dev).bfloat16()
res2 = {}
ms = timeit(lambda: DN.ln_modulate(xs, sh, sc)); res2["ln_modulate"] = (ms, 2 * S * 3072 * 2)
norm = torch.nn.LayerNorm(3072, elementwise_affine=False, eps=1e-6).to(dev)
def eager_ln():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return (norm(xs) * (1 + sc.unsqueeze(1)) + sh.unsqueeze(1)).to(torch.bfloat16)
ms = timeit(eager_ln); res2["  eager ATen chain"] = (ms, 2 * S * 3072 * 2)
ys = torch.randn(1, S, 3072, device=dev).bfloat16()
ms = timeit(lambda: DN.gate_residual(xs, ys, sc)); res2["gate_residual"] = (ms, 3 * S * 3072 * 2)
ms = timeit(lambda: xs + ys * sc.unsqueeze(1)); res2["  eager ATen chain"+" "] = (ms, 3 * S * 3072 * 2)
lin1 = torch.randn(1, S, 9216 + 12288, device=dev).bfloat16()
cat = torch.empty(1, S, 3072 + 12288, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: DN.gelu_tanh_into(lin1[:, :, 9216:], cat[:, :, 3072:])); res2["gelu_tanh_into"] = (ms, 2 * S * 12288 * 2)
act = torch.nn.GELU(approximate="tanh")
att = torch.randn(1, S, 3072, device=dev).bfloat16()
ms = timeit(lambda: torch.cat((att, act(lin1[:, :, 9216:])), 2)); res2["  eager gelu + cat"] = (ms, 2 * S * 12288 * 2)
ms = timeit(lambda: SC.residual_apply(xs, ys)); res2["residual_apply"] = (ms, 3 * S * 3072 * 2)
ms = timeit(lambda: SC.residual_store(xs, ys)); res2["residual_store"] = (ms, 3 * S * 3072 * 2)
ms = timeit(lambda: gilbert.mapping_tensors_device(32, 45, 80)); res2["gilbert tables (device)"] = (ms, 0)
dl2h, _ = gilbert.mapping_tensors_device(32, 45, 80)
ms = timeit(lambda: gilbert.block_neighbor_bits_device(32, 45, 80, dl2h)); res2["adjacency bits (device)"] = (ms, 0)
import time
t0 = time.perf_counter(); gilbert.mapping_tensors(32, 45, 80); gilbert.block_neighbor_mapping(32, 45, 80); t1 = time.perf_counter()
print(f"host walker tables + adjacency: {(t1 - t0) * 1e3:.1f} ms")
for k_, (ms, b) in res2.items():
    print(f"{k_:26s} {ms:8.3f} ms  {b/ms/1e6:9.1f} GB/s  frac_of_hbm_peak {b/ms/1e6/peak:5.2f}" if b else f"{k_:26s} {ms:8.3f} ms")

# ---- Wan-14B shapes: the WanRMSNorm + fp64-RoPE prologue and the fp32 -> bf16 cast fused with the block pooling
from jenga_b200 import wan as WAN
Lw, Hw = 75600, 40
xw = torch.randn(1, Lw, Hw * 128, device=dev)
ww = torch.ones(Hw * 128, device=dev)
freqs = WAN.rope_freqs(128)
_, h2lw = gilbert.mapping_tensors(21, 45, 80, sliced=True)
h2lw = h2lw.to(dev)
res3 = {}
import os
xb = xw.bfloat16(); wb = ww.bfloat16()
for mode in ("vector", "vector64", "scalar"):
    os.environ["JENGA_WAN_PROLOGUE"] = mode
    ms = timeit(lambda: WAN.norm_rope(xw, ww, Hw, (21, 45, 80), freqs, h2lw)); res3[f"wan_prologue {mode} (f32 in)"] = (ms, Lw * Hw * 128 * (4 + 2))
    ms = timeit(lambda: WAN.norm_rope(xb, wb, Hw, (21, 45, 80), freqs, h2lw)); res3[f"wan_prologue {mode} (bf16 in)"] = (ms, Lw * Hw * 128 * (2 + 2))
os.environ.pop("JENGA_WAN_PROLOGUE", None)
q32 = torch.randn(1, Lw, Hw, 128, device=dev)
castbuf = torch.empty(q32.shape, dtype=torch.bfloat16, device=dev)
nbw = (Lw + 127) // 128
ms = timeit(lambda: block_pool(q32, nbw, cast_out=castbuf)); res3["block_pool f32->bf16 + cast"] = (ms, Lw * Hw * 128 * (4 + 2))
for k_, (ms, b) in res3.items():
    print(f"{k_:34s} {ms:8.3f} ms  {b/ms/1e6:9.1f} GB/s  frac_of_hbm_peak {b/ms/1e6/peak:5.2f}")
