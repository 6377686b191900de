"""a-13 / f-2 on the GPU: residual cache ops (bit-exact vs ATen), the TeaCache gate against a
restatement of jenga_wan.py:597-626 in torch+numpy, the HunyuanVideo skip schedule driving the
cache, and CUDA-graph replay of the hot path (bit-identical to eager)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n", [8 * 1000 + 5, 3 * 4096 * 64])
def test_residual_ops_are_bit_exact_vs_aten(dtype, n):
    from jenga_b200 import stepcache as S
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(n, generator=g, device="cuda").to(dtype)
    b = torch.randn(n, generator=g, device="cuda").to(dtype)
    res = S.residual_store(a, b)
    assert torch.equal(res, a - b)                       # jenga_hyvideo.py:179
    x = b.clone()
    want = b.clone()
    want += res                                          # jenga_hyvideo.py:130
    S.residual_apply(x, res)
    assert torch.equal(x, want)
    buf = torch.empty_like(a)
    assert S.residual_store(a, b, out=buf) is buf and torch.equal(buf, a - b)


def _reference_gate(seq, coeffs, thresh, ret_steps, cutoff_steps):
    """jenga_wan.py:597-626 restated (one parity stream): returns (decisions, accumulated values)."""
    acc, prev, out, accs = 0.0, None, [], []
    f = np.poly1d(coeffs)
    for cnt, x in seq:
        if cnt < ret_steps or cnt >= cutoff_steps or prev is None:
            calc, acc = True, 0.0
        else:
            acc += f(((x - prev).abs().mean() / prev.abs().mean()).cpu().item())
            if acc < thresh:
                calc = False
            else:
                calc, acc = True, 0.0
        prev = x.clone()
        out.append(calc)
        accs.append(acc)
    return out, accs


def test_teacache_gate_matches_reference_state_machine(dtype=torch.float32):
    from jenga_b200.stepcache import TeaCache
    coeffs = [-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02]  # jenga_wan.py:1087
    steps, ret, cut, thresh = 30, 10, 60, 0.15
    tc = TeaCache(coeffs, thresh, ret, cut)
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(1, 6, 1536, generator=g, device="cuda")
    seqs = {0: [], 1: []}
    got = {0: [], 1: []}
    accs = {0: [], 1: []}
    x = {0: base.clone(), 1: base.clone() * 0.9}
    for cnt in range(2 * steps):
        par = cnt % 2
        x[par] = x[par] + (0.004 + 0.002 * (cnt % 7)) * torch.randn(x[par].shape, generator=g, device="cuda")
        cur = x[par].to(dtype)
        seqs[par].append((cnt, cur.clone()))
        assert tc.cnt == cnt
        got[par].append(tc.gate(cur))
        accs[par].append(tc.accumulated(par))
        tc.step()
    for par in (0, 1):
        want, want_acc = _reference_gate(seqs[par], coeffs, thresh, ret, cut)
        assert got[par] == want, (par, got[par], want)
        assert np.allclose(accs[par], want_acc, rtol=1e-4, atol=1e-6)
        assert 3 < sum(want) < len(want)                  # the sequence exercises both outcomes
    assert torch.equal(tc._prev[1], seqs[1][-1][1])       # prev <- cur (the reference's .clone())


def test_step_skip_cache_follows_the_reference_schedule():
    """jenga_hyvideo.py:120-179 with synthetic "blocks" (x -> x*1.01+0.5): 23 computed steps of 50,
    forced compute at the stage switch, residual replay on skipped steps == a torch restatement."""
    from jenga_b200.stepcache import NON_SKIP_STEPS, StepSkipCache
    cache = StepSkipCache(50)
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn(1, 2048, 256, generator=g, device="cuda").bfloat16()
    ref_img, prev_res, computed = img.clone(), None, []
    for step in range(50):
        if step == 26:
            cache.start_stage = True                       # pipeline_hunyuan_video_prores.py:763-767
        noise = torch.randn(img.shape, generator=g, device="cuda").bfloat16() * 0.1
        img, ref_img = img + noise, ref_img + noise
        calc_ref = step in NON_SKIP_STEPS or step == 26
        if cache.should_calc():
            ori = img
            img = img * 1.01 + 0.5
            cache.store(img, ori)
            computed.append(step)
        else:
            cache.apply(img)
        if calc_ref:
            ori = ref_img.clone()
            ref_img = ref_img * 1.01 + 0.5
            prev_res = ref_img - ori
        else:
            ref_img += prev_res
        cache.step()
        assert torch.equal(img, ref_img), step
    assert computed == sorted(set(NON_SKIP_STEPS) | {26}) and len(computed) == 23
    assert cache.cnt == 0                                  # wrapped (:221-223)


def test_cuda_graph_replay_of_the_hot_path_is_bit_identical():
    """f-2: 6 consecutive block hot paths (prologue -> selection -> carved attention) captured in
    one CUDA graph; replay on fresh inputs == eager result, bit for bit."""
    import bench
    from jenga_b200 import hyvideo as HY
    from jenga_b200.stepcache import GraphedHotPath
    dev = torch.device("cuda", 0)
    wl = bench.workload("tiny", 0.7)
    inp = bench.build_inputs(wl, dev, heads=4)
    L, T, H = inp["n_img"], wl["text_tokens"], 4
    g = torch.Generator(device=dev).manual_seed(9)
    img_qkv = torch.randn(1, L, 3 * H * 128, generator=g, device=dev).bfloat16()
    txt_qkv = torch.randn(1, T, 3 * H * 128, generator=g, device=dev).bfloat16()
    ws = [(1 + 0.1 * torch.randn(128, generator=g, device=dev)).bfloat16() for _ in range(4)]
    ang = torch.rand(L, 64, generator=g, device=dev) * 6.28
    freqs = (torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous())
    nbr, cu = inp["nbr"], inp["cu"]

    def loop(iq, tq):
        outs = []
        for _ in range(6):
            q, k, v, pools = HY.attention_prologue(iq, tq, H, *ws, eps=1e-6, freqs_cis=freqs)
            outs.append(HY.carved_attention_from_pools(q, k, v, pools, top_k=inp["top_k"], text_blocks=2,
                                                       text_amp=0.0, block_neighbor_list=nbr, p_remain_rates=0.3,
                                                       cu_seqlens_q=cu))
        return outs

    si, st = img_qkv.clone(), txt_qkv.clone()
    gh = GraphedHotPath(loop, si, st)
    new_i = torch.randn(img_qkv.shape, generator=g, device=dev).bfloat16()
    new_t = torch.randn(txt_qkv.shape, generator=g, device=dev).bfloat16()
    eager = loop(new_i, new_t)
    torch.cuda.synchronize()
    gh.copy_inputs(new_i, new_t)
    outs = gh.replay()
    torch.cuda.synchronize()
    assert len(outs) == 6 and all(torch.equal(a, b) for a, b in zip(outs, eager))


@pytest.mark.parametrize("shape,size", [((1, 16, 8, 17, 30), (8, 34, 60)), ((1, 16, 32, 33, 60), (32, 45, 80)),
                                        ((2, 4, 5, 6, 7), (5, 12, 14))])
def test_prores_stage_switch_matches_the_reference_chain(shape, size):
    """pipeline_hunyuan_video_prores.py:721-731 restated with the reference's own torch ops:
    predict_x0_from_xt -> F.interpolate(trilinear) -> add_noise_to_step, all fp32."""
    from jenga_b200.stepcache import prores_switch
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    lat = torch.randn(shape, generator=g, device="cuda").bfloat16()
    pred = torch.randn(shape, generator=g, device="cuda").bfloat16()
    noise = torch.randn(shape[:2] + size, generator=g, device="cuda").bfloat16()
    sig_i, sig_last, sig_next = 0.8421, 0.0, 0.8107
    d_sigma = sig_last - sig_i
    x0 = lat.to(torch.float32) + pred.to(torch.float32) * d_sigma
    up = torch.nn.functional.interpolate(x0, size=size, mode="trilinear")
    ref = up.to(torch.float32) * (1.0 - sig_next) + noise.to(torch.float32) * sig_next
    got = prores_switch(lat, pred, noise, size, d_sigma, sig_next)
    assert got.dtype == torch.float32 and got.shape == ref.shape
    err = (got - ref).abs().max().item()
    print(f"\n[prores switch {shape}->{size}] max |diff| {err:.2e}")
    assert err <= 2e-6 * max(1.0, ref.abs().max().item())

