"""Wan2.1-side glue of the hot path (wan/modules/model_mul.py:134-180 WanSelfAttention.forward):
fused WanRMSNorm + RoPE prologue, the block-count rules and the operator call."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import check, lib
from .attention import _require_cuda, _stream_ptr, block_sparse_attention_variant


def rope_freqs(head_dim: int = 128, max_len: int = 1024, theta: float = 10000.0) -> torch.Tensor:
    """The complex128 table WanModel builds (model_mul.py:502-507 with rope_params :29-37)."""
    def params(dim):
        f = torch.outer(torch.arange(max_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        return torch.polar(torch.ones_like(f), f)
    d = head_dim
    return torch.cat([params(d - 4 * (d // 6)), params(2 * (d // 6)), params(2 * (d // 6))], dim=1)


def block_counts(seq_len: int, sa_drop_rate: float, per_block_tokens: int = 128):
    """model_mul.py:161-164: (selected_top_k, first_frame_blocks)."""
    num_blocks = math.ceil(seq_len / per_block_tokens)
    return math.ceil(int(num_blocks * (1 - sa_drop_rate))), math.ceil(num_blocks // 21)


_FREQ_CACHE: dict = {}


def _device_freqs(freqs: torch.Tensor, dev) -> torch.Tensor:
    """The complex128 table as doubles [rows, 64, 2] on `dev`, and its fp32 (hi, lo) split [rows, 64, 4].  WanModel keeps `self.freqs` on the CPU
    (model_mul.py:502-507) and the reference moves it every forward; here the device copy is made once
    per table (keyed by the tensor object, dropped with it)."""
    import weakref
    key = (id(freqs), str(dev))
    hit = _FREQ_CACHE.get(key)
    if hit is not None and hit[0]() is freqs and hit[2] == freqs._version:
        return hit[1]
    fr = torch.view_as_real(freqs.to(torch.complex128)).to(dev).contiguous()
    hi = fr.to(torch.float32)
    lo = (fr - hi.to(torch.float64)).to(torch.float32)
    hilo = torch.cat([hi, lo], dim=-1).contiguous()          # [rows, 64, 4] = re_hi, im_hi, re_lo, im_lo
    _FREQ_CACHE[key] = (weakref.ref(freqs, lambda _r, k=key: _FREQ_CACHE.pop(k, None)), (fr, hilo), freqs._version)
    return fr, hilo


def norm_rope(x: torch.Tensor, weight: torch.Tensor | None, heads: int, grid_size, freqs: torch.Tensor,
              freq_remap: torch.Tensor | None = None, eps: float = 1e-6) -> torch.Tensor:
    """rope_apply(norm(x).view(b, s, n, d), grid_sizes, freqs, freq_remap) as the carved
    attention consumes it: [B, L, H, 128] bf16.  x: [B, L, H*128] bf16 or fp32."""
    _require_cuda(x)
    B, L, Cc = x.shape
    if Cc != heads * 128 or x.stride(2) != 1:
        raise ValueError("x must be [B, L, heads*128] with contiguous channels")
    dev = x.device
    out = torch.empty((B, L, heads, 128), dtype=torch.bfloat16, device=dev)
    a = _lib.JengaWanPrologueArgs()
    a.x = x.data_ptr()
    code = {torch.bfloat16: _lib.JENGA_BF16, torch.float32: _lib.JENGA_F32}
    if x.dtype not in code:
        raise ValueError("x must be bf16 or fp32")
    a.x_dtype = code[x.dtype]
    if weight is not None:
        weight = weight.to(dev).contiguous()
        if weight.dtype not in code:
            raise ValueError("weight must be bf16 or fp32")
        a.w, a.w_dtype = weight.data_ptr(), code[weight.dtype]
    else:
        a.w, a.w_dtype = None, _lib.JENGA_F32
    a.batch, a.heads, a.head_dim, a.tokens = B, heads, 128, L
    a.stride_b, a.stride_s, a.eps = x.stride(0), x.stride(1), eps
    fr, hilo = _device_freqs(freqs, dev)
    a.freqs, a.freq_rows = fr.data_ptr(), fr.shape[0]
    a.freqs_hilo = hilo.data_ptr()
    a.grid_f, a.grid_h, a.grid_w = (int(v) for v in grid_size)
    if freq_remap is not None:
        freq_remap = freq_remap.to(device=dev, dtype=torch.int64).contiguous()
        a.freq_remap = freq_remap.data_ptr()
    else:
        a.freq_remap = None
    a.out = out.data_ptr()
    with torch.cuda.device(dev):
        check(lib.jenga_wan_prologue(C.byref(a), _stream_ptr(dev)), "wan_prologue")
    return out


def self_attention(q_lin, k_lin, v, norm_q_w, norm_k_w, heads, grid_size, freqs, sa_drop_rate, *,
                   p_remain_rates=0.8, freq_remap=None, block_neighbor_list=None, eps=1e-6):
    """WanSelfAttention.forward's attention part for sa_drop_rate > 0.25 (model_mul.py:160-176):
    q_lin/k_lin [B, L, C] are self.q(x)/self.k(x), v is self.v(x).view(b, s, n, d)."""
    q = norm_rope(q_lin, norm_q_w, heads, grid_size, freqs, freq_remap, eps)
    k = norm_rope(k_lin, norm_k_w, heads, grid_size, freqs, freq_remap, eps)
    top_k, ff = block_counts(q.shape[1], sa_drop_rate)
    return block_sparse_attention_variant("wan", q, k, v, top_k, text_blocks=0,
                                          block_neighbor_list=block_neighbor_list,
                                          p_remain_rates=p_remain_rates, first_frame_blocks=ff,
                                          shape_xfuse=True)
