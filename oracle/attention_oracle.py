"""CPU restatement of the reference's AttenCarve operator — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the product: nothing under jenga_b200/ imports it.

Parity status: PINNED for the attention kernel semantics (a-9) against the unmodified
reference Triton kernel executed with TRITON_INTERPRET=1 in the build container (fp16
fixtures, tests/golden/make_golden.py); PINNED for the mask builder (a-8) against the
unmodified reference builder run on CPU with CUDA-autocast dtype flow emulated; the text rows
(a-10) restate flash_attn_func's published semantics (softmax(q k^T * scale) v, fp32
accumulate) because FlashAttention-2 (flash-attn 2.x, un-vendored CUDA) cannot run on CPU.

Every function cites the reference lines it follows; paths are relative to the Jenga repo.
All arithmetic is done in float32 with explicit re-rounding where the reference rounds.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.44269504  # literal used by the reference launcher (…triton_diffres.py:172)


def _round_like(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return x.to(dtype).to(torch.float32)


def carved_attention_rows(
    q: torch.Tensor,  # [B, H, Nq, D]  (Nq multiple of 128)
    k: torch.Tensor,  # [B, H, S, D]   (S  multiple of 128)
    v: torch.Tensor,
    onehot: torch.Tensor,  # [B, H, Nq/128, S/128] bool
    seqlen: int,
    sm_scale: float,
    text_amp: float = 0.0,
    text_block_start: int = 0,
    block: int = 128,
) -> torch.Tensor:
    """hyvideo/modules/attention_block_triton_diffres.py:38-136 (kernel) + :139-196 (launcher).

    Follows the kernel tile by tile: Q pre-scaled by sm_scale*log2e in fp32 and re-rounded to
    the input dtype (:87-88); per live key block (ascending, :94-97): s = q~ k^T with fp32
    accumulation (:110), + text_amp for key blocks >= text_block_start (:113-114), -inf for
    key columns >= seqlen (:117-118); online softmax in base 2 (:121-132) with P rounded to
    the input dtype before the PV product (:128) and l accumulated from the unrounded p (:131);
    O = acc / l (:135); rows >= seqlen stay zero (:136 with zeros_like :156); q blocks starting
    at or past seqlen are skipped entirely (:59-61).
    """
    dt = q.dtype
    B, H, Nq, D = q.shape
    S = k.shape[2]
    nq, nb = Nq // block, S // block
    assert onehot.shape == (B, H, nq, nb), (onehot.shape, (B, H, nq, nb))
    qk_scale = torch.tensor(sm_scale * LOG2E, dtype=torch.float32)
    qt = _round_like(q.float() * qk_scale, dt)
    kf, vf = k.float(), v.float()
    out = torch.zeros(B, H, Nq, D, dtype=torch.float32)
    cols_in_block = torch.arange(block)
    for b in range(B):
        for h in range(H):
            for m in range(nq):
                if m * block >= seqlen:
                    continue
                qm = qt[b, h, m * block:(m + 1) * block]
                m_i = torch.full((block,), float("-inf"))
                l_i = torch.zeros(block)
                acc = torch.zeros(block, D)
                for j in torch.nonzero(onehot[b, h, m]).flatten().tolist():
                    kj = kf[b, h, j * block:(j + 1) * block]
                    vj = vf[b, h, j * block:(j + 1) * block]
                    s = qm @ kj.T
                    if j >= text_block_start:
                        s = s + text_amp
                    s = torch.where((j * block + cols_in_block)[None, :] < seqlen, s,
                                    torch.tensor(float("-inf")))
                    m_new = torch.maximum(m_i, s.max(dim=1).values)
                    alpha = torch.exp2(m_i - m_new)
                    p = torch.exp2(s - m_new[:, None])
                    acc = acc * alpha[:, None] + _round_like(p, dt) @ vj
                    l_i = l_i * alpha + p.sum(dim=1)
                    m_i = m_new
                o = acc / l_i[:, None]
                rows = m * block + torch.arange(block)
                o = torch.where((rows < seqlen)[:, None], o, torch.zeros_like(o))
                out[b, h, m * block:(m + 1) * block] = o
    return out.to(dt)


def dense_attention_rows(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sm_scale: float,
                         kv_limit: int | None = None) -> torch.Tensor:
    """Text rows: flash_attn_func(q, k, v, causal=False, softmax_scale=sm_scale)
    (…triton_diffres.py:371-380; FlashAttention-2 published forward: S = q k^T in fp32,
    P = softmax(S * scale) in fp32, P rounded to the input dtype for P v, fp32 accumulate).
    Layout here is [B, H, N, D].  No key mask: padded text keys ARE attended (SURVEY A-3);
    kv_limit is only for callers that emulate an unpadded key length."""
    dt = q.dtype
    s = (q.float() @ k.float().transpose(-1, -2)) * sm_scale
    if kv_limit is not None and kv_limit < k.shape[2]:
        s[..., kv_limit:] = float("-inf")
    m = s.max(dim=-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    o = (_round_like(p, dt) @ v.float()) / l
    return o.to(dt)


def build_block_onehot(
    query: torch.Tensor,  # [B, H, Nq_img, D]  (image rows only; already in the MMA dtype)
    key: torch.Tensor,    # [B, H, S, D]       (all rows, incl. text / padding)
    top_k: int,
    text_start_block: int,
    num_blocks: int,
    prob_threshold: float,
    text_blocks: int,
    block_neighbor_list: torch.Tensor | None = None,
    first_frame_blocks: int = 0,
    block: int = 128,
    tie_break: str = "low",
    return_probs: bool = False,
):
    """…triton_diffres.py:198-295 (HunyuanVideo) and wan/…:306-411 (adds first_frame_blocks)
    with the dtype flow the reference sees under torch.autocast(cuda, bfloat16)
    (SURVEY Appendix A-5): mean -> input dtype; bmm -> input dtype; * D^-1/2 -> input dtype;
    softmax / sort / cumsum in fp32.
    bf16 scores tie often.  The reference's torch.sort(descending=True) is not stable: on CUDA
    (segmented radix sort) equal keys keep index order -> tie_break="low", which is what the
    product kernel implements; on CPU the order of equal keys is arbitrary (tie_break="high"
    is the other extreme), so fixtures are compared modulo swaps inside one tie group."""
    dt = query.dtype
    B, H, Nq, D = query.shape
    nq = (Nq + block - 1) // block
    qp = _round_like(query.float().reshape(B, H, -1, block, D).mean(dim=-2), dt)   # :216
    kp = _round_like(key.float().reshape(B, H, -1, block, D).mean(dim=-2), dt)     # :217
    scores = _round_like(qp @ kp.transpose(-1, -2), dt)                            # :227 bmm
    # `bf16_tensor * python_float`: fp32 op-math with the scalar kept in fp32, one rounding
    scores = _round_like(scores * torch.tensor(D ** -0.5, dtype=torch.float32), dt)
    normal = scores[..., :text_start_block]                                        # :235
    probs = torch.softmax(normal, dim=-1)                                          # :238 fp32
    if tie_break == "low":
        sorted_probs, indices = torch.sort(probs, dim=-1, descending=True, stable=True)  # :241
    else:
        sorted_probs, indices = torch.sort(probs.flip(-1), dim=-1, descending=True, stable=True)
        indices = probs.shape[-1] - 1 - indices
    csum = torch.cumsum(sorted_probs, dim=-1)                                      # :242
    need = (csum <= prob_threshold).sum(dim=-1) + 1                                # :245-246
    need = torch.clamp(need, min=top_k)                                            # :247-250
    need = torch.clamp(need, max=indices.shape[-1])
    onehot = torch.zeros(B, H, nq, num_blocks, dtype=torch.bool)
    rank = torch.arange(indices.shape[-1]).view(1, 1, 1, -1)
    keep = rank < need.unsqueeze(-1)                                               # :264
    sel = torch.zeros(B, H, nq, indices.shape[-1], dtype=torch.bool)
    sel.scatter_(-1, indices, keep)                                                # :268-276
    onehot[..., :text_start_block] = sel
    if block_neighbor_list is not None:                                            # :280-289
        nm = block_neighbor_list[:nq, :text_start_block].bool()
        onehot[:, :, :nm.shape[0], :nm.shape[1]] |= nm[None, None]
    if first_frame_blocks > 0:                                                     # wan/…:400-406
        onehot[:, :, :first_frame_blocks, :first_frame_blocks] = True
    if text_blocks > 0:                                                            # :292-293
        onehot[..., text_start_block:min(text_start_block + text_blocks, num_blocks)] = True
    return (onehot, probs) if return_probs else onehot


def block_sparse_attention(
    query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, top_k: int,
    block_size_M: int = 128, block_size_N: int = 128, cu_seqlens_q=None, cu_seqlens_kv=None,
    max_seqlen_q=None, max_seqlen_kv=None, text_blocks: int = 2, text_amp: float = 0.0,
    block_neighbor_list=None, shape_xfuse: bool = False, p_remain_rates: float = 0.5,
    first_frame_blocks: int = 0, variant: str = "hyvideo", return_mask: bool = False,
    mask_override: torch.Tensor | None = None,
):
    """The whole operator, [B,S,H,D] in -> [B,S,H*D] out.
    variant "hyvideo": …/hyvideo/modules/attention_block_triton_diffres.py:298-424
    variant "i2v"    : …/hyvideo_i2v/…:298-423 (pads S to x128, trims, text_blocks=4 default)
    variant "wan"    : …/wan/…:413-561 (always pads, bf16 inside, text_blocks=0, first frame)"""
    assert block_size_M == 128 and block_size_N == 128
    out_dtype = query.dtype
    q = query.transpose(1, 2)
    k = key.transpose(1, 2)
    v = value.transpose(1, 2)
    B, H, S, D = q.shape
    use_cu = cu_seqlens_q is not None and cu_seqlens_kv is not None and variant != "wan"
    pad = (128 - S % 128) % 128
    if use_cu:
        seqlen = int(cu_seqlens_q[1])                                              # :328-329
    else:
        seqlen = S
    if variant == "hyvideo":
        pad = 0  # the HunyuanVideo variant never pads (:327-335); the I2V one does (:323-328)
    if pad:
        q = torch.nn.functional.pad(q, [0, 0, 0, pad])
        k = torch.nn.functional.pad(k, [0, 0, 0, pad])
        v = torch.nn.functional.pad(v, [0, 0, 0, pad])
    if variant == "wan":
        q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)  # wan/…:456-463
    Sp = q.shape[2]
    sm_scale = D ** -0.5
    num_blocks = (Sp + 127) // 128
    normal_blocks = num_blocks - text_blocks
    normal_tokens = normal_blocks * 128
    outs = []
    mask = None
    if normal_blocks > 0:
        qn = q[:, :, :normal_tokens]
        mask = mask_override if mask_override is not None else build_block_onehot(
            qn, k, top_k, normal_blocks, num_blocks, p_remain_rates, text_blocks,
            block_neighbor_list, first_frame_blocks)
        outs.append(carved_attention_rows(qn, k, v, mask, seqlen, sm_scale, text_amp,
                                          normal_blocks))
    if text_blocks > 0:
        outs.append(dense_attention_rows(q[:, :, normal_tokens:], k, v, sm_scale))
    o = torch.cat(outs, dim=2)[:, :, :S]
    o = o.permute(0, 2, 1, 3)
    if not shape_xfuse:
        o = o.reshape(B, S, -1)
    o = o.to(out_dtype) if variant == "wan" else o
    return (o, mask) if return_mask else o
