"""Full-size (BASELINE.json configs[1]: HunyuanVideo 720x1280x125f, S=115456, H=24) checks through
size-independent properties, plus sampled rows against an fp32 torch evaluation on the GPU."""
import math
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.fixture(scope="module")
def hy720p():
    import bench
    from jenga_b200.attention import bits_to_onehot
    dev = torch.device("cuda", 0)
    wl = bench.workload("hy720p", 0.7)
    inp = bench.build_inputs(wl, dev)
    out, bits = bench.run_operator(wl, inp, return_bits=True)
    torch.cuda.synchronize()
    return wl, inp, out, bits


def test_mask_invariants_full_size(hy720p):
    from jenga_b200.attention import bits_to_onehot, neighbour_bits
    wl, inp, out, bits = hy720p
    nb_img, nb = inp["nb_img"], inp["nb_img"] + wl["text_blocks"]
    assert bits.shape == (1, 24, nb_img, (nb + 31) // 32)
    oh = bits_to_onehot(bits, bits.shape[-1] * 32)
    assert not oh[..., nb:].any()                       # no bits past the key blocks
    assert oh[..., nb_img:nb].all()                     # text columns always on (ref :292-293)
    counts = oh[..., :nb_img].sum(-1)
    assert int(counts.min()) >= inp["top_k"]            # top_k is a floor (ref :247-250)
    nbr = inp["nbr"].to(oh.device)
    assert (oh[0, :, :, :nb_img] | ~nbr[None]).all()    # neighbour rows are a subset (ref :280-289)
    assert int(counts.max()) <= nb_img


def test_sampled_rows_against_fp32_torch(hy720p):
    """Rows of 6 query blocks x 3 heads recomputed in fp32 from the same mask, with the
    reference's rounding points (Q pre-scale re-rounding, P in bf16, text rows dense)."""
    from jenga_b200.attention import bits_to_onehot
    wl, inp, out, bits = hy720p
    q, k, v = inp["q"], inp["k"], inp["v"]
    S, nb_img = inp["S"], inp["nb_img"]
    nb = nb_img + wl["text_blocks"]
    seqlen = inp["n_img"] + wl["text_valid"]
    oh = bits_to_onehot(bits, nb)
    qk_scale = torch.tensor((128 ** -0.5) * 1.44269504, dtype=torch.float32, device=q.device)
    o4 = out.view(1, S, 24, 128)
    worst = 0.0
    for h in (0, 11, 23):
        kh, vh = k[0, :, h].float(), v[0, :, h].float()
        for m in (0, 1, 450, 899, 900, 901):
            rows = slice(m * 128, (m + 1) * 128)
            qm = q[0, rows, h].float()
            if m < nb_img:
                live = oh[0, h, m].nonzero().flatten()
                cols = (live[:, None] * 128 + torch.arange(128, device=q.device)[None]).flatten()
                qt = (qm * qk_scale).to(torch.bfloat16).float()
                s = qt @ kh[cols].T
                s = s.masked_fill(cols[None, :] >= seqlen, float("-inf"))
                p = torch.exp2(s - s.max(-1, keepdim=True).values)
            else:  # text rows: dense, FlashAttention semantics, padded text keys attended
                cols = torch.arange(S, device=q.device)
                s = (qm @ kh.T) * (128 ** -0.5)
                p = torch.exp(s - s.max(-1, keepdim=True).values)
            ref = (p.to(torch.bfloat16).float() @ vh[cols]) / p.sum(-1, keepdim=True)
            got = o4[0, rows, h].float()
            rms = ref.pow(2).mean().sqrt()
            d = (got - ref).abs()
            assert (d <= 2e-2 * rms + 2.0 ** -7 * ref.abs()).all(), (h, m, (d / rms).max().item())
            assert d.mean() <= 3e-3 * rms  # see test_attn_gpu.assert_close for the noise floor
            worst = max(worst, (d / rms).max().item())
    print("worst max|err|/rms over sampled rows:", worst)


def test_value_linearity_is_bit_exact(hy720p):
    """O is linear in V and scaling by 2 is exact in every format on the path, so
    attention(q, k, 2v) must equal 2*attention(q, k, v) bit for bit at full size."""
    import bench
    wl, inp, out, bits = hy720p
    out2 = bench.run_operator(wl, inp, v=inp["v"] * 2)
    torch.cuda.synchronize()
    assert torch.equal(out2, out * 2)
    assert torch.isfinite(out.float()).all()


def test_gather_scatter_round_trip_full_size():
    """[1, 115200, 3072] bf16 hidden states along the 32x45x80 curve and back (a-4)."""
    from jenga_b200 import gilbert
    from jenga_b200.hyvideo import gather_tokens
    dev = torch.device("cuda", 0)
    l2h, h2l = gilbert.mapping_tensors(32, 45, 80)
    x = torch.randn(1, 115200, 3072, device=dev, dtype=torch.bfloat16)
    g = gather_tokens(x, h2l.to(dev))
    idx = torch.tensor([0, 1, 57599, 115199], device=dev)
    assert torch.equal(g[0, idx], x[0, h2l.to(dev)[idx]])
    back = gather_tokens(g, l2h.to(dev))
    assert torch.equal(back, x)
