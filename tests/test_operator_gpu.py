"""End-to-end parity of the drop-in operator (reference signature) on the GPU."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import synth  # noqa: E402

from test_attn_gpu import assert_close  # noqa: E402


def _case(variant, dtype):
    if variant == "hyvideo":
        H, n_img, T = 3, 12, 256
        S = n_img * 128 + T
        kw = dict(top_k=3, text_blocks=2, text_amp=0.431, p_remain_rates=0.3)
        cu = torch.tensor([0, n_img * 128 + 180, S], dtype=torch.int32)
    elif variant == "hyvideo_i2v":
        H, n_img, T = 2, 10, 400
        S = n_img * 128 + T  # padded to 14 blocks inside
        kw = dict(top_k=3, text_blocks=4, text_amp=0.0, p_remain_rates=0.3)
        cu = torch.tensor([0, n_img * 128 + 300, S], dtype=torch.int32)
    else:
        H, n_img = 2, 16
        S = n_img * 128 - 40
        kw = dict(top_k=8, text_blocks=0, text_amp=0.0, p_remain_rates=0.9, first_frame_blocks=2)
        cu = None
    nbk = (S + 127) // 128
    q = synth.peaky(H, nbk, 128, 2.0, 301)[:, :, :S].transpose(1, 2).contiguous()
    k = synth.peaky(H, nbk, 128, 2.0, 302)[:, :, :S].transpose(1, 2).contiguous()
    v = synth.normal((1, S, H, 128), 303)
    if variant == "wan":
        q, k, v = q.float(), k.float(), v.to(torch.bfloat16)  # wan: fp32 q,k + bf16 v
    else:
        q, k, v = q.to(dtype), k.to(dtype), v.to(dtype)
    nbr = synth.band_neighbours(n_img)
    return q, k, v, cu, nbr, kw, n_img


@pytest.mark.parametrize("variant,dtype", [("hyvideo", torch.bfloat16), ("hyvideo", torch.float16),
                                           ("hyvideo_i2v", torch.bfloat16), ("wan", torch.bfloat16)])
def test_operator_matches_oracle(variant, dtype):
    from jenga_b200.attention import bits_to_onehot, block_sparse_attention_variant
    from oracle import attention_oracle as orc
    q, k, v, cu, nbr, kw, n_img = _case(variant, dtype)
    dev = "cuda"
    cu_d = cu.to(dev) if cu is not None else None
    out, bits = block_sparse_attention_variant(
        variant, q.to(dev), k.to(dev), v.to(dev), cu_seqlens_q=cu_d, cu_seqlens_kv=cu_d,
        block_neighbor_list=nbr, return_mask_bits=True, **kw)
    torch.cuda.synchronize()
    S = q.shape[1]
    nb = (S + 127) // 128
    assert out.shape == (1, S, q.shape[2] * 128)
    assert out.dtype == q.dtype  # wan returns the query dtype (fp32)
    mask = bits_to_onehot(bits, nb).cpu()
    okw = dict(kw)
    ref = orc.block_sparse_attention(q, k, v, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                     block_neighbor_list=nbr, variant="i2v" if variant == "hyvideo_i2v" else variant,
                                     mask_override=mask, **okw)
    assert_close(out.cpu(), ref, torch.bfloat16 if variant == "wan" else dtype)
    # and the mask itself: oracle selection on the same inputs (CUDA tie order)
    ref_mask = orc.block_sparse_attention(q, k, v, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                          block_neighbor_list=nbr,
                                          variant="i2v" if variant == "hyvideo_i2v" else variant,
                                          return_mask=True, **okw)[1]
    agree = (mask == ref_mask).float().mean().item()
    assert agree >= 0.995, agree


def test_shape_xfuse_and_no_text():
    from jenga_b200.attention import block_sparse_attention_variant
    q, k, v, cu, nbr, kw, n_img = _case("wan", torch.bfloat16)
    dev = "cuda"
    o = block_sparse_attention_variant("wan", q.to(dev), k.to(dev), v.to(dev), shape_xfuse=True,
                                       block_neighbor_list=nbr, **kw)
    assert o.shape == q.shape and torch.isfinite(o).all()


@pytest.mark.parametrize("name", [n for n, *_ in synth.OPERATOR_CASES])
def test_operator_matches_reference_golden(name):
    """Product (through the C-ABI) against the output of the UNMODIFIED reference operator on the
    same fp16 inputs (tests/golden/operator_fp16.npz; Triton kernel run in the interpreter, the
    FlashAttention-2 text-row call replaced by fp32 attention — see make_golden.py).  Tolerance:
    the fp16 attention bound of test_attn_gpu.assert_close."""
    import numpy as np
    from jenga_b200.attention import block_sparse_attention_variant
    gold = np.load(HERE / "golden" / "operator_fp16.npz")
    c = synth.operator_case(name)
    dev = "cuda"
    out = block_sparse_attention_variant(
        c["variant"], c["q"].to(dev), c["k"].to(dev), c["v"].to(dev), top_k=c["top_k"],
        cu_seqlens_q=c["cu"].to(dev), cu_seqlens_kv=c["cu"].to(dev), text_blocks=c["text_blocks"],
        text_amp=c["amp"], block_neighbor_list=c["nbr"], shape_xfuse=c["xfuse"], p_remain_rates=c["p"])
    torch.cuda.synchronize()
    ref = torch.from_numpy(gold[name + "/o"])
    assert out.shape == ref.shape and out.dtype == ref.dtype
    assert_close(out.cpu(), ref, torch.float16)
