"""GPU parity of block pooling + selection (a-8) through the C-ABI."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
import synth  # noqa: E402

from test_oracle_cpu import assert_masks_equal_modulo_ties  # noqa: E402


def _run_select(c, dev="cuda"):
    from jenga_b200.attention import (bits_to_onehot, block_pool, mask_onehot_to_bits,
                                      select_blocks)
    n_img, n_txt = c["n_img"], c["n_txt"]
    nb = n_img + n_txt
    q = c["q"].transpose(1, 2).contiguous().to(dev)  # [1,S,H,D]
    k = c["k"].transpose(1, 2).contiguous().to(dev)
    qp = block_pool(q[:, :n_img * 128], n_img)
    kp = block_pool(k, n_img)
    nbr = mask_onehot_to_bits(c["nbr"].to(dev)) if c["nbr"] is not None else None
    bits, counts = select_blocks(qp, kp, n_img=n_img, nb=nb, top_k=c["top_k"], p_threshold=c["p"],
                                 text_blocks=n_txt, first_frame_blocks=c["ff"], nbr_bits=nbr,
                                 return_counts=True)
    torch.cuda.synchronize()
    return bits_to_onehot(bits, nb).cpu(), counts.cpu(), qp.cpu(), kp.cpu()


@pytest.mark.parametrize("name", [n for n, *_ in synth.MASK_CASES])
def test_selection_matches_reference_builder_golden(name):
    from oracle import attention_oracle as orc
    gold = np.load(HERE / "golden" / "mask_builder.npz")
    c = synth.mask_case(name)
    nb = c["n_img"] + c["n_txt"]
    got, counts, qp, kp = _run_select(c)
    ref = torch.from_numpy(gold[name + "/mask"])
    low, probs = orc.build_block_onehot(c["q"][:, :, :c["n_img"] * 128], c["k"], c["top_k"],
                                        c["n_img"], nb, c["p"], c["n_txt"], c["nbr"], c["ff"],
                                        tie_break="low", return_probs=True)
    # pooled means: same rounding point as the reference (bf16 of an fp32 mean)
    qp_ref = c["q"][:, :, :c["n_img"] * 128].float().reshape(1, c["H"], -1, 128, 128).mean(-2)
    assert (qp.float() - qp_ref).abs().max() <= 2.0 ** -8 * qp_ref.abs().max() + 1e-6
    # vs the reference's own output: identical modulo swaps inside a tie group
    assert_masks_equal_modulo_ties(got, ref, probs, c["n_img"])
    # vs the oracle with CUDA tie order: exact
    assert (got == low).all(), int((got != low).sum())


def test_selection_large_random_vs_oracle():
    """256+2 blocks, 4 heads: per-row counts equal and sets equal up to fp32 summation-order
    effects at the cumulative-probability cut (Jaccard >= 0.99)."""
    from jenga_b200.attention import bits_to_onehot, block_pool, select_blocks
    from oracle import attention_oracle as orc
    H, n_img, n_txt, top_k, p = 4, 256, 2, 64, 0.3
    nb = n_img + n_txt
    q = synth.peaky(H, nb, 128, 2.0, 991).bfloat16()
    k = synth.peaky(H, nb, 128, 2.0, 992).bfloat16()
    dev = "cuda"
    qd = q.transpose(1, 2).contiguous().to(dev)
    kd = k.transpose(1, 2).contiguous().to(dev)
    bits = select_blocks(block_pool(qd[:, :n_img * 128], n_img), block_pool(kd, n_img), n_img=n_img,
                         nb=nb, top_k=top_k, p_threshold=p, text_blocks=n_txt)
    got = bits_to_onehot(bits, nb).cpu()
    ref = orc.build_block_onehot(q[:, :, :n_img * 128], k, top_k, n_img, nb, p, n_txt)
    inter = (got & ref).sum(-1).float()
    union = (got | ref).sum(-1).float()
    assert (inter / union).mean() >= 0.99
    assert (got.sum(-1) == ref.sum(-1)).float().mean() >= 0.98
    assert (got[..., n_img:] == ref[..., n_img:]).all()


@pytest.mark.parametrize("n_img,n_txt,top_k,p,first", [(900, 2, 270, 0.3, 0), (256, 0, 128, 0.9, 12),
                                                     (37, 4, 5, 0.5, 0), (1024, 2, 1, 0.05, 0)])
def test_two_kernel_path_is_bit_identical_to_fused_kernel(n_img, n_txt, top_k, p, first):
    """With a score workspace the selection runs as GEMM + one-warp-per-row kernel; without it as
    one fused kernel.  Same arithmetic in the same order: bit rows and counts must be identical,
    including rows with exact probability ties at the cut (duplicated key blocks)."""
    from jenga_b200.attention import select_blocks
    H, nb = 3, n_img + n_txt
    g = torch.Generator().manual_seed(n_img)
    qp = (torch.randn(1, H, n_img, 128, generator=g) * 1.5).bfloat16().cuda()
    kp = (torch.randn(1, H, nb, 128, generator=g) * 1.5).bfloat16().cuda()
    kp[:, :, 1:n_img:3] = kp[:, :, 0:n_img - 1:3][:, :, :kp[:, :, 1:n_img:3].shape[2]]  # duplicated blocks -> ties
    nbr = (torch.randint(0, 2**31 - 1, (n_img, (nb + 31) // 32), generator=g, dtype=torch.int64) &
           torch.randint(0, 2**31 - 1, (n_img, (nb + 31) // 32), generator=g, dtype=torch.int64)).int().cuda()
    kw = dict(n_img=n_img, nb=nb, top_k=top_k, p_threshold=p, text_blocks=n_txt, first_frame_blocks=first,
              nbr_bits=nbr, return_counts=True)
    b1, c1 = select_blocks(qp, kp, use_workspace=True, **kw)
    b0, c0 = select_blocks(qp, kp, use_workspace=False, **kw)
    assert torch.equal(c0, c1)
    assert torch.equal(b0, b1), int((b0 != b1).sum())
