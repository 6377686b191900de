"""GPU tests of the FlashAttention shims (dense mode of the kernel) and of the fused
prologue -> selection -> attention sequence a DiT block would run."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import synth  # noqa: E402

from test_attn_gpu import assert_close  # noqa: E402


def _sdpa(q, k, v, scale):
    # [B,S,H,D] fp32 reference on the GPU
    o = torch.nn.functional.scaled_dot_product_attention(
        q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), scale=scale)
    return o.transpose(1, 2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_flash_attn_func_matches_sdpa(dtype):
    from jenga_b200.flash_attn_shim import flash_attn_func
    dev = "cuda"
    q = synth.normal((2, 300, 3, 128), 1).to(dtype).to(dev)   # ragged: 300 rows, 3 q blocks
    k = synth.normal((2, 1000, 3, 128), 2).to(dtype).to(dev)  # ragged keys: masked tail
    v = synth.normal((2, 1000, 3, 128), 3).to(dtype).to(dev)
    out = flash_attn_func(q, k, v, softmax_scale=0.1)
    ref = _sdpa(q, k, v, 0.1)
    assert out.shape == q.shape and out.dtype == dtype
    assert_close(out, ref, dtype)
    with pytest.raises(NotImplementedError):
        flash_attn_func(q, k, v, causal=True)


def test_flash_attn_varlen_matches_reference_call_pattern():
    """hyvideo/modules/attenion.py:109-117 with cu_seqlens = [0, valid, S] (get_cu_seqlens :34-57):
    tokens [0,valid) attend each other, the padding tokens [valid,S) attend among themselves."""
    from jenga_b200.flash_attn_shim import flash_attn_varlen_func
    dev = "cuda"
    S, valid, H = 640, 500, 2
    q = synth.normal((S, H, 128), 4).bfloat16().to(dev)
    k = synth.normal((S, H, 128), 5).bfloat16().to(dev)
    v = synth.normal((S, H, 128), 6).bfloat16().to(dev)
    cu = torch.tensor([0, valid, S], dtype=torch.int32, device=dev)
    out = flash_attn_varlen_func(q, k, v, cu, cu, S, S)
    ref = torch.cat([_sdpa(q[None, :valid], k[None, :valid], v[None, :valid], 128 ** -0.5),
                     _sdpa(q[None, valid:], k[None, valid:], v[None, valid:], 128 ** -0.5)], dim=1)[0]
    assert_close(out, ref, torch.bfloat16)


def test_fused_block_sequence_equals_operator():
    """attention_prologue -> carved_attention_from_pools (3 launches) must equal
    block_sparse_attention on the prologue's q,k,v (pool + pool + select + attention):
    identical masks (same pooled values) and identical attention output."""
    from jenga_b200.attention import block_sparse_attention_variant
    from jenga_b200.hyvideo import attention_prologue, carved_attention_from_pools, select_block_num
    dev = "cuda"
    t, h, w = 4, 8, 16  # 512 image tokens
    L, T, H = t * h * w, 256, 3
    img = (1.5 * synth.normal((1, L, 3 * H * 128), 51)).bfloat16().to(dev)
    txt = (0.7 * synth.normal((1, T, 3 * H * 128), 52)).bfloat16().to(dev)
    ws = [(1.0 + 0.2 * synth.normal((128,), 53 + i)).bfloat16() for i in range(4)]
    cos = torch.cos(synth.normal((L, 128), 60)).float()
    sin = torch.sin(synth.normal((L, 128), 60)).float()
    q, k, v, pools = attention_prologue(img, txt, H, *ws, eps=1e-6, freqs_cis=(cos, sin))
    nbr = synth.band_neighbours(L // 128)
    cu = torch.tensor([0, L + 180, L + T], dtype=torch.int32, device=dev)
    top_k = select_block_num(0.5, L)
    a = carved_attention_from_pools(q, k, v, pools, top_k=top_k, text_blocks=2, text_amp=0.2,
                                    block_neighbor_list=nbr, p_remain_rates=0.3, cu_seqlens_q=cu)
    b = block_sparse_attention_variant("hyvideo", q, k, v, top_k, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                       text_blocks=2, text_amp=0.2, block_neighbor_list=nbr,
                                       p_remain_rates=0.3)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert select_block_num(0.75, 115200) == 225 and select_block_num(0.8, 115200) == 179
    assert select_block_num(0.85, 14400, world_size=8) == 128   # SP rounding quirk


def test_host_pipeline_equals_device_operator():
    """Head-group pipelined host->device->host path must reproduce the device-resident operator
    bit for bit (per-head independence)."""
    from jenga_b200.attention import block_sparse_attention_variant
    from jenga_b200.host_pipeline import HostPipelinedAttention
    dev = "cuda"
    H, nbi, T = 6, 8, 256
    S = nbi * 128 + T
    nb = S // 128
    q = synth.peaky(H, nb, 128, 2.0, 901).transpose(1, 2).contiguous().bfloat16()
    k = synth.peaky(H, nb, 128, 2.0, 902).transpose(1, 2).contiguous().bfloat16()
    v = synth.normal((1, S, H, 128), 903).bfloat16()
    nbr = synth.band_neighbours(nbi)
    cu = torch.tensor([0, nbi * 128 + 180, S], dtype=torch.int32, device=dev)
    kw = dict(cu_seqlens_q=cu, cu_seqlens_kv=cu, text_blocks=2, text_amp=0.25, block_neighbor_list=nbr,
              p_remain_rates=0.3)
    ref = block_sparse_attention_variant("hyvideo", q.to(dev), k.to(dev), v.to(dev), 3, shape_xfuse=True, **kw)
    pipe = HostPipelinedAttention(1, S, H, 128, torch.bfloat16, dev, groups=3)
    hq, hk, hv = q.pin_memory(), k.pin_memory(), v.pin_memory()
    hout = torch.empty_like(hq).pin_memory()
    for _ in range(2):  # twice: exercises buffer reuse across calls
        hout.zero_()
        pipe(hq, hk, hv, hout, 3, **kw)
        torch.cuda.synchronize()
        assert torch.equal(hout, ref.cpu())


def test_wan_self_attention_composition():
    """wan.self_attention == norm_rope(q), norm_rope(k) -> block_sparse_attention (wan variant)."""
    from jenga_b200 import wan
    from jenga_b200.attention import block_sparse_attention_variant
    dev = "cuda"
    f, h, w = 4, 8, 16   # 512 tokens = 4 blocks
    L, H = f * h * w, 2
    xq = synth.normal((1, L, H * 128), 71).bfloat16().to(dev)
    xk = synth.normal((1, L, H * 128), 72).bfloat16().to(dev)
    v = synth.normal((1, L, H, 128), 73).bfloat16().to(dev)
    wq = (1 + 0.1 * synth.normal((H * 128,), 74)).to(dev)
    wk = (1 + 0.1 * synth.normal((H * 128,), 75)).to(dev)
    freqs = wan.rope_freqs(128)
    nbr = synth.band_neighbours(4)
    out = wan.self_attention(xq, xk, v, wq, wk, H, (f, h, w), freqs, 0.5, p_remain_rates=0.8, block_neighbor_list=nbr)
    q = wan.norm_rope(xq, wq, H, (f, h, w), freqs)
    k = wan.norm_rope(xk, wk, H, (f, h, w), freqs)
    top_k, ff = wan.block_counts(L, 0.5)
    ref = block_sparse_attention_variant("wan", q, k, v, top_k, text_blocks=0, block_neighbor_list=nbr,
                                         p_remain_rates=0.8, first_frame_blocks=ff, shape_xfuse=True)
    torch.cuda.synchronize()
    assert out.shape == (1, L, H, 128) and torch.equal(out, ref)


def test_dit_loop_harness_runs_the_schedule():
    from jenga_b200 import dit_loop, gilbert
    dev = "cuda"
    t, h, w = 2, 8, 16  # 256 image tokens
    L, T, H = t * h * w, 256, 2
    img = synth.normal((1, L, 3 * H * 128), 81).bfloat16().to(dev)
    txt = synth.normal((1, T, 3 * H * 128), 82).bfloat16().to(dev)
    ws = [torch.ones(128, dtype=torch.bfloat16, device=dev) for _ in range(4)]
    cos = torch.cos(synth.normal((L, 128), 83)).to(dev)
    sin = torch.sin(synth.normal((L, 128), 83)).to(dev)
    nbr = gilbert.gilbert_block_neighbor_mapping(t, h, w)
    cu = torch.tensor([0, L + 180, L + T], dtype=torch.int32, device=dev)
    sec, calls = dit_loop.run_hot_path_loop(img, txt, H, ws, (cos, sin), nbr, cu, max_computed_steps=2)
    assert calls == 2 * 60 and sec > 0
