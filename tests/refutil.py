"""Shared helpers of the reference-backed GPU parity tests (and bench.py's gpu_reference leg).

fp64 ground truth of the operator's MATH (no intermediate rounding) on a sample of query blocks,
error statistics relative to it, and the call sequence of the unmodified reference operator
(oracle/ref_loader.py).  Test infrastructure only.
"""
from __future__ import annotations

import math

import torch

BLOCK = 128
LN2 = math.log(2.0)


def pad_rows(x: torch.Tensor, rows: int) -> torch.Tensor:
    """[1,S,H,D] -> [1,rows,H,D] zero padded (what the I2V / Wan variants do, wan/…:448-451)."""
    if x.shape[1] == rows:
        return x
    return torch.nn.functional.pad(x, [0, 0, 0, 0, 0, rows - x.shape[1]])


@torch.no_grad()
def fp64_truth(q, k, v, onehot, qblocks, *, n_img_blocks, seqlen, text_amp=0.0, sm_scale=None):
    """Exact softmax attention in fp64 for the query blocks `qblocks` of every head.

    q,k,v: [1,S,H,D] device tensors (any float dtype; values are taken as they are),
    onehot: [1,H,n_img_blocks,nb] bool — the block mask of the image rows.
    Image rows (block < n_img_blocks): keys = mask-selected blocks, key columns >= seqlen hidden,
    text key blocks (>= n_img_blocks) biased by text_amp in log2 units
    (…triton_diffres.py:96-118).  Text rows: dense over ALL nb*128 keys incl. padding, no bias
    (:371-380).  Returns {qb: [H,128,D] float64}."""
    _, S, H, D = q.shape
    nb = onehot.shape[-1]
    Sp = nb * BLOCK
    sm_scale = D ** -0.5 if sm_scale is None else sm_scale
    kd = pad_rows(k, Sp)[0].transpose(0, 1).double()          # [H,Sp,D]
    vd = pad_rows(v, Sp)[0].transpose(0, 1).double()
    qd = pad_rows(q, Sp)[0].transpose(0, 1)
    cols = torch.arange(Sp, device=q.device)
    out = {}
    for qb in qblocks:
        Q = qd[:, qb * BLOCK:(qb + 1) * BLOCK].double()        # [H,128,D]
        s = torch.matmul(Q, kd.transpose(1, 2)) * sm_scale     # [H,128,Sp]
        if qb < n_img_blocks:
            live = onehot[0, :, qb].repeat_interleave(BLOCK, dim=-1)            # [H,Sp]
            if text_amp != 0.0:
                s[:, :, n_img_blocks * BLOCK:] += text_amp * LN2
            dead = (~live)[:, None, :] | (cols >= seqlen)[None, None, :]
            s = s.masked_fill(dead, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[qb] = torch.matmul(p, vd)
        del s, p
    return out


def error_stats(got, truth: dict, *, seqlen_rows=None):
    """got: [1,S,H*D] or [1,S,H,D]; truth from fp64_truth.  Returns (max|e|/rms, mean|e|/rms, rms)
    over the sampled blocks (rows >= seqlen_rows of a block are skipped: both implementations
    leave them zero / undefined)."""
    any_t = next(iter(truth.values()))
    H, _, D = any_t.shape
    g = got.reshape(1, got.shape[1], H, D)
    num_max, num_sum, cnt, sq = 0.0, 0.0, 0, 0.0
    for qb, t in truth.items():
        r0 = qb * BLOCK
        r1 = min(r0 + BLOCK, g.shape[1])
        if seqlen_rows is not None:
            r1 = min(r1, seqlen_rows)
        if r1 <= r0:
            continue
        x = g[0, r0:r1].transpose(0, 1).double()
        tt = t[:, : r1 - r0]
        e = (x - tt).abs()
        num_max = max(num_max, e.max().item())
        num_sum += e.sum().item()
        cnt += e.numel()
        sq += tt.pow(2).sum().item()
    rms = math.sqrt(sq / max(cnt, 1)) + 1e-30
    return num_max / rms, num_sum / max(cnt, 1) / rms, rms


def pair_stats(a, b):
    """max|a-b|/rms(b), mean|a-b|/rms(b) over two same-shaped tensors."""
    a, b = a.double(), b.double()
    rms = b.pow(2).mean().sqrt().item() + 1e-30
    d = (a - b).abs()
    return d.max().item() / rms, d.mean().item() / rms


def reference_call(op, variant, q, k, v, *, top_k, cu, text_blocks, text_amp, nbr, p_remain,
                   first_frame=0, shape_xfuse=False):
    """The unmodified reference operator under the pipelines' autocast
    (pipeline_hunyuan_video_prores.py:663-665, jenga_wan.py:135)."""
    kw = dict(cu_seqlens_q=cu, cu_seqlens_kv=cu, text_blocks=text_blocks, text_amp=text_amp,
              block_neighbor_list=nbr, shape_xfuse=shape_xfuse, p_remain_rates=p_remain)
    if variant == "wan":
        kw["first_frame_blocks"] = first_frame
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return op.block_sparse_attention(q, k, v, top_k, **kw)


def reference_mask(op, variant, q, k, *, top_k, text_blocks, nbr, p_remain, first_frame=0):
    """The reference builder's one-hot mask for the same inputs, called the way
    block_sparse_attention_combined calls it (…triton_diffres.py:345-356, wan :476-487)."""
    B, S, H, D = q.shape
    nb = (S + BLOCK - 1) // BLOCK
    qh = pad_rows(q, nb * BLOCK).transpose(1, 2)
    kh = pad_rows(k, nb * BLOCK).transpose(1, 2)
    if variant == "wan":
        qh, kh = qh.to(torch.bfloat16), kh.to(torch.bfloat16)
    n_img = nb - text_blocks
    kw = dict(text_start_block=n_img, num_blocks=nb, prob_threshold=p_remain, text_blocks=text_blocks,
              block_neighbor_list=nbr)
    if variant == "wan":
        kw["first_frame_blocks"] = first_frame
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return op._build_block_index_with_importance_optimized(qh[:, :, :n_img * BLOCK], kh, top_k, BLOCK, BLOCK, **kw)


def mask_agreement(got: torch.Tensor, ref: torch.Tensor, n_img: int):
    """(fraction of identical rows, mean Jaccard, fraction of equal counts, text columns exact)."""
    same_rows = (got == ref).all(-1).float().mean().item()
    inter = (got & ref).sum(-1).float()
    union = (got | ref).sum(-1).float().clamp_min(1)
    jac = (inter / union).mean().item()
    counts = (got.sum(-1) == ref.sum(-1)).float().mean().item()
    text_ok = bool((got[..., n_img:] == ref[..., n_img:]).all().item())
    return same_rows, jac, counts, text_ok
