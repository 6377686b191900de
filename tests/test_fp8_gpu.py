"""f-3 (opt-in): FP8 (e4m3) P.V variant of the carved attention.  NOT the reference's arithmetic: P and
V lose 5 mantissa bits on the way into the second product, so it has its OWN tolerance row, derived
the same way as the bf16 one (fp64 ground truth on the same inputs): relative rounding error of an
e4m3 element is up to 2^-4 (rms ~2^-4/sqrt(3) = 3.6 %), uncorrelated across keys, so the output error
is ~3.6 % of the output RMS (rms) for P plus the same again for V: mean |err| <= 6e-2 * RMS,
max <= 0.35 * RMS.  Q.K^T, the softmax, l and the mask are untouched (identical selection)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
import refutil  # noqa: E402


def test_v_quantisation_round_trips_within_e4m3_precision():
    from jenga_b200.attention import quantize_v_fp8
    g = torch.Generator(device="cuda").manual_seed(1)
    v = (torch.randn(1, 1000, 3, 128, generator=g, device="cuda") * torch.tensor([0.5, 1.0, 3.0], device="cuda")[None, None, :, None]).bfloat16()
    v8, amax = quantize_v_fp8(v)
    assert v8.dtype == torch.uint8 and v8.shape == v.shape
    want = v.float().abs().amax(dim=(0, 1, 3))
    assert torch.equal(amax, want)
    back = v8.view(torch.float8_e4m3fn).float() * (amax / 448.0)[None, None, :, None]
    rel = (back - v.float()).abs() / v.float().abs().clamp_min(1e-3 * want[None, None, :, None])
    assert rel.max().item() <= 2.0 ** -4 + 1e-3            # e4m3: 3 mantissa bits, round to nearest
    # ATen's own cast gives the same bytes
    ref8 = (v.float() * (448.0 / amax)[None, None, :, None]).to(torch.float8_e4m3fn).view(torch.uint8)
    assert (ref8 == v8).float().mean().item() >= 0.999


@pytest.mark.parametrize("case", ["tiny", "hy_turbo_s0"])
def test_fp8_pv_variant_against_fp64_truth(case):
    import bench
    from jenga_b200.attention import bits_to_onehot, block_sparse_attention_variant
    wl = bench.workload(case, 0.7)
    inp = bench.build_inputs(wl, torch.device("cuda", 0), heads=2, seed=99)
    kw = dict(cu_seqlens_q=inp["cu"], cu_seqlens_kv=inp["cu"], text_blocks=wl["text_blocks"], text_amp=wl["text_amp"],
              block_neighbor_list=inp["nbr"], p_remain_rates=wl["p_remain"], return_mask_bits=True)
    ref, bits = block_sparse_attention_variant(wl["variant"], inp["q"], inp["k"], inp["v"], inp["top_k"], **kw)
    got, bits8 = block_sparse_attention_variant(wl["variant"], inp["q"], inp["k"], inp["v"], inp["top_k"], pv_fp8=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(bits, bits8)                                     # selection is untouched
    S = inp["S"]
    nb = (S + 127) // 128
    n_img = nb - wl["text_blocks"]
    mask = bits_to_onehot(bits, nb)
    seqlen = int(inp["cu"][1].item())
    blocks = sorted({0, n_img // 2, n_img - 1}) + [n_img]
    truth = refutil.fp64_truth(inp["q"], inp["k"], inp["v"], mask, blocks, n_img_blocks=n_img, seqlen=seqlen,
                               text_amp=wl["text_amp"])
    img = {b: t for b, t in truth.items() if b < n_img}
    b_max, b_mean, _ = refutil.error_stats(ref, img, seqlen_rows=seqlen)
    f_max, f_mean, _ = refutil.error_stats(got, img, seqlen_rows=seqlen)
    t_max, t_mean, _ = refutil.error_stats(got, {n_img: truth[n_img]})
    print(f"\n[fp8 P.V {case}] vs fp64: bf16 path max {b_max:.2e} mean {b_mean:.2e}; fp8 path max {f_max:.2e} mean {f_mean:.2e}; "
          f"text rows fp8 {t_max:.2e} / {t_mean:.2e}")
    assert f_mean <= 6e-2 and f_max <= 0.35
    assert t_mean <= 6e-2 and t_max <= 0.35
    assert f_mean >= 2 * b_mean                                         # it really is the lower-precision path
