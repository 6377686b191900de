"""Fused elementwise chains of the DiT blocks around the attention (first slice).

Host mirror of csrc/dense_fused.cu; each function replaces a chain of ATen kernels of
hyvideo/modules/models_mul_block_gc_ha_multigpu.py (file:line in each docstring) with one pass
that keeps the chain's rounding points under torch.autocast(bf16).  bf16 CUDA tensors only; the
callers (jenga_b200/blocks.py) fall back to the reference's own ops for anything else."""
from __future__ import annotations

import torch

from ._lib import check, lib
from .attention import _require_cuda, _stream_ptr


def _rows2d(x: torch.Tensor):
    """[..., C] tensor whose leading dims collapse to uniformly strided rows -> (rows, row_stride)."""
    if x.stride(-1) != 1:
        raise ValueError("channel dim must be contiguous")
    C = x.shape[-1]
    rows = x.numel() // C
    if x.dim() == 1:
        return 1, C
    stride = x.stride(-2)
    # leading dims must continue the same row pitch (true for [B, L, C] slices of contiguous tensors)
    expect = stride * x.shape[-2]
    for d in range(x.dim() - 3, -1, -1):
        if x.shape[d] != 1 and x.stride(d) != expect:
            raise ValueError("rows are not uniformly strided")
        expect *= x.shape[d]
    return rows, stride


def ln_modulate_supported(x: torch.Tensor, norm: torch.nn.Module, shift, scale) -> bool:
    return (isinstance(norm, torch.nn.LayerNorm) and norm.weight is None and norm.bias is None
            and x.is_cuda and x.dtype == torch.bfloat16 and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16 and x.dim() == 3 and x.shape[0] == 1
            and shift is not None and scale is not None and shift.dtype == torch.bfloat16
            and scale.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0 and x.shape[-1] <= 6144
            and x.stride(-1) == 1 and x.stride(-2) % 8 == 0)


def ln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """`modulate(LayerNorm(x), shift=shift, scale=scale)` followed by the bf16 cast the next Linear
    applies under autocast (models_mul…:196-199,297-304,409; modulate_layers.py:31-49).
    x [1, L, C] bf16, shift/scale [1, C] bf16 -> [1, L, C] bf16."""
    _require_cuda(x, shift, scale)
    rows, xs = _rows2d(x)
    C = x.shape[-1]
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    sc, sh = scale.reshape(-1).contiguous(), shift.reshape(-1).contiguous()
    if sc.numel() != C or sh.numel() != C:
        raise ValueError("shift / scale must hold one value per channel (batch 1)")
    with torch.cuda.device(x.device):
        check(lib.jenga_ln_modulate(x.data_ptr(), xs, sc.data_ptr(), sh.data_ptr(), out.data_ptr(), C, rows, C,
                                    float(eps), _stream_ptr(x.device)), "ln_modulate")
    return out


def gate_residual_supported(x, y, gate) -> bool:
    return (x.is_cuda and x.dtype == y.dtype == torch.bfloat16 and gate is not None and gate.dtype == torch.bfloat16
            and x.shape == y.shape and x.dim() == 3 and x.shape[0] == 1 and x.shape[-1] % 8 == 0
            and x.stride(-1) == 1 and y.stride(-1) == 1 and x.stride(-2) % 8 == 0 and y.stride(-2) % 8 == 0)


def gate_residual(x: torch.Tensor, y: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """`x + apply_gate(y, gate=gate)` (models_mul…:295-315,:500; modulate_layers.py:52-68)."""
    _require_cuda(x, y, gate)
    rows, xs = _rows2d(x)
    _, ys = _rows2d(y)
    C = x.shape[-1]
    g = gate.reshape(-1).contiguous()
    if g.numel() != C:
        raise ValueError("gate must hold one value per channel (batch 1)")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.jenga_gate_residual(x.data_ptr(), xs, y.data_ptr(), ys, g.data_ptr(), out.data_ptr(), C, rows, C,
                                      _stream_ptr(x.device)), "gate_residual")
    return out


def gelu_tanh_into(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """`out[...] = nn.GELU(approximate="tanh")(x)` for strided [1, L, C] bf16 views (the MLP half of
    linear1's output into the concatenation buffer, models_mul…:413,:499)."""
    _require_cuda(x, out)
    if x.shape != out.shape or x.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise ValueError("x / out must be bf16 tensors of one shape")
    rows, xs = _rows2d(x)
    _, os_ = _rows2d(out)
    with torch.cuda.device(x.device):
        check(lib.jenga_gelu_tanh(x.data_ptr(), xs, out.data_ptr(), os_, rows, x.shape[-1], _stream_ptr(x.device)),
              "gelu_tanh")
    return out
