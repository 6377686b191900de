"""World-size-2 gloo test of the Ulysses wrapper's layout logic (no GPU): the sequence-parallel
result must equal the single-process operator on the full tensors, bit for bit, because the
selection and the attention are per-head and the exchange is pure data movement."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE / "golden"))


def _inputs():
    import synth
    H, n_img_blocks, T = 4, 4, 256
    n_img = n_img_blocks * 128
    nb = n_img_blocks + T // 128
    q = synth.peaky(H, nb, 128, 2.0, 801).transpose(1, 2).contiguous().bfloat16()  # [1,S,H,D]
    k = synth.peaky(H, nb, 128, 2.0, 802).transpose(1, 2).contiguous().bfloat16()
    v = synth.normal((1, nb * 128, H, 128), 803).bfloat16()
    nbr = synth.band_neighbours(n_img_blocks)
    return q, k, v, nbr, n_img, T


def _oracle_attn(q, k, v, **kw):
    from oracle import attention_oracle as orc
    kw.pop("block_size_M", None)
    kw.pop("block_size_N", None)
    return orc.block_sparse_attention(q, k, v, variant="hyvideo", **kw)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jenga_b200.ulysses import UlyssesCarvedAttention, my_parallel_attention, seq_to_heads, heads_to_seq
        q, k, v, nbr, n_img, T = _inputs()
        n_loc = n_img // world
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        # layout round trip
        x = q[:, sl]
        y = seq_to_heads(x)
        assert y.shape == (1, n_img, q.shape[2] // world, 128)
        h = q.shape[2] // world
        assert torch.equal(y, q[:, :n_img, rank * h:(rank + 1) * h])
        assert torch.equal(heads_to_seq(y), x)
        ql = torch.cat([q[:, sl], q[:, n_img:]], 1)
        kl = torch.cat([k[:, sl], k[:, n_img:]], 1)
        vl = torch.cat([v[:, sl], v[:, n_img:]], 1)
        cu = torch.tensor([0, n_loc + 180, n_loc + T], dtype=torch.int32)
        sp = UlyssesCarvedAttention(attn_fn=_oracle_attn)
        out = my_parallel_attention(sp, ql, kl, vl, n_loc, n_loc, cu, cu, top_k=2, text_amp=0.3,
                                    block_neighbor_list=nbr, p_remain_rates=0.3)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_ulysses_equals_single_process_operator():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    q, k, v, nbr, n_img, T = _inputs()
    cu = torch.tensor([0, n_img + 180, n_img + T], dtype=torch.int32)
    full = _oracle_attn(q, k, v, top_k=2, cu_seqlens_q=cu, cu_seqlens_kv=cu, text_amp=0.3,
                        block_neighbor_list=nbr, p_remain_rates=0.3, text_blocks=T // 128)
    n_loc = n_img // world
    for r in range(world):
        got = ret[r]
        assert got.shape == (1, n_loc + T, q.shape[2] * 128)
        assert torch.equal(got[:, :n_loc], full[:, r * n_loc:(r + 1) * n_loc])
        assert torch.equal(got[:, n_loc:], full[:, n_img:])
