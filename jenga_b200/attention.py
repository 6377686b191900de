"""Host-side mirror of the reference's AttenCarve operator for the B200 path.

`block_sparse_attention` keeps the reference signature
(hyvideo/modules/attention_block_triton_diffres.py:399-424, hyvideo_i2v twin :398-423,
wan twin :535-562) and argument meaning; all arithmetic runs in libjenga_b200.so.
torch is used for device memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import JengaAttnArgs, JengaError, check, lib

BLOCK = 128


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.JENGA_BF16
    if t.dtype == torch.float16:
        return _lib.JENGA_F16
    raise ValueError(f"unsupported dtype {t.dtype}: the carved attention runs in bf16 or fp16")


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise JengaError("jenga_b200 has no CPU path: tensors must live on a CUDA device")


def mask_onehot_to_bits(onehot: torch.Tensor) -> torch.Tensor:
    """[..., nq, nb] bool one-hot block mask (the tensor the reference builder returns,
    …triton_diffres.py:252-295) -> [..., nq, ceil(nb/32)] int32 bit rows."""
    _require_cuda(onehot)
    oh = onehot.contiguous()
    if oh.dtype != torch.bool and oh.dtype != torch.uint8:
        raise ValueError("one-hot mask must be bool or uint8")
    nb = oh.shape[-1]
    words = (nb + 31) // 32
    rows = oh.numel() // nb
    bits = torch.empty(*oh.shape[:-1], words, dtype=torch.int32, device=oh.device)
    with torch.cuda.device(oh.device):
        check(lib.jenga_mask_onehot_to_bits(oh.data_ptr(), bits.data_ptr(), rows, nb, words,
                                            _stream_ptr(oh.device)), "mask_onehot_to_bits")
    return bits


def carved_attention_fwd(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,  # [B, S, H, D], D contiguous
    mask_bits: torch.Tensor | None,  # [B, H, nq_sparse, words] int32
    nq_sparse: int, nq_dense: int,
    sm_scale: float, text_amp: float = 0.0, text_block_start: int = 1 << 30,
    kv_limit_sparse: int | None = None, q_limit_sparse: int | None = None,
    kv_limit_dense: int | None = None, out: torch.Tensor | None = None,
    err_flag: torch.Tensor | None = None,
) -> torch.Tensor:
    """One launch of the sm_100a carved-attention kernel (C-ABI jenga_carved_attn_fwd).
    Sparse q-blocks [0, nq_sparse) follow the reference Triton kernel (:38-136); dense
    q-blocks [nq_sparse, nq_sparse+nq_dense) follow its flash_attn_func text rows (:371-380).
    Returns `out` [B, Sq, H, D]."""
    _require_cuda(q, k, v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be [B, S, H, D]")
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    if k.shape != (B, Sk, H, D) or v.shape != (B, Sk, H, D):
        raise ValueError("k / v shape mismatch")
    if not (q.dtype == k.dtype == v.dtype):
        raise ValueError("q, k, v must share a dtype")
    for t in (q, k, v):
        if t.stride(3) != 1:
            raise ValueError("head_dim must be contiguous")
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
    a = JengaAttnArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.dtype = _dtype_code(q)
    a.batch, a.heads, a.head_dim = B, H, D
    a.q_rows, a.kv_rows = Sq, Sk
    a.q_stride_b, a.q_stride_s, a.q_stride_h = q.stride(0), q.stride(1), q.stride(2)
    a.k_stride_b, a.k_stride_s, a.k_stride_h = k.stride(0), k.stride(1), k.stride(2)
    a.v_stride_b, a.v_stride_s, a.v_stride_h = v.stride(0), v.stride(1), v.stride(2)
    a.o_stride_b, a.o_stride_s, a.o_stride_h = out.stride(0), out.stride(1), out.stride(2)
    a.nq_sparse, a.nq_dense = nq_sparse, nq_dense
    if nq_sparse > 0:
        if mask_bits is None:
            raise ValueError("mask_bits required for sparse query blocks")
        _require_cuda(mask_bits)
        if mask_bits.dtype != torch.int32 or not mask_bits.is_contiguous():
            raise ValueError("mask_bits must be contiguous int32")
        if tuple(mask_bits.shape[:3]) != (B, H, nq_sparse):
            raise ValueError(f"mask_bits shape {tuple(mask_bits.shape)} != ({B},{H},{nq_sparse},W)")
        a.mask_bits = mask_bits.data_ptr()
        a.mask_words = mask_bits.shape[3]
    else:
        a.mask_bits, a.mask_words = None, 0
    a.sm_scale, a.text_amp, a.text_block_start = sm_scale, text_amp, min(text_block_start, 1 << 30)
    a.kv_limit_sparse = Sk if kv_limit_sparse is None else kv_limit_sparse
    a.q_limit_sparse = Sq if q_limit_sparse is None else q_limit_sparse
    a.kv_limit_dense = Sk if kv_limit_dense is None else kv_limit_dense
    a.err_flag = err_flag.data_ptr() if err_flag is not None else None
    with torch.cuda.device(q.device):
        check(lib.jenga_carved_attn_fwd(C.byref(a), _stream_ptr(q.device)), "carved_attn_fwd")
    return out
