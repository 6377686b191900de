"""Deterministic synthetic inputs shared by the golden generator and the tests.

Integer-only generation (splitmix64 over the flat index + an Irwin-Hall sum of four uniforms)
so that every machine and library version reproduces the same float32 values bit for bit; the
committed fixtures then only need to hold the reference's OUTPUTS.
"""
import numpy as np
import torch

_M = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(_M)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(_M)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(_M)
    return z ^ (z >> np.uint64(31))


def normal(shape, seed: int) -> torch.Tensor:
    """Approximately N(0,1) float32 tensor, a pure function of (shape, seed)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) * np.uint64(4) + np.uint64(seed) * np.uint64(0x100000001B3)
        acc = np.zeros(n, dtype=np.int64)
        for j in range(4):
            acc += (_splitmix64(idx + np.uint64(j)) >> np.uint64(40)).astype(np.int64)  # 24-bit uniforms
    # sum of four U[0,2^24): mean 2^25, variance 4 * 2^48 / 12
    x = (acc - (1 << 25)).astype(np.float64) / float(1 << 24) * np.sqrt(3.0)
    return torch.from_numpy(x.astype(np.float32).reshape(shape))


def peaky(H: int, nb: int, D: int, gain: float, seed: int, block: int = 128) -> torch.Tensor:
    """Model-like tokens (SURVEY §8d regime P): per-block centroids, smoothed along the curve,
    plus within-block noise.  Returns [1, H, nb*block, D] float32."""
    c = normal((H, nb, D), seed)
    pad = torch.nn.functional.pad(c.transpose(1, 2), (2, 2), mode="replicate")
    c = torch.nn.functional.avg_pool1d(pad, 5, 1).transpose(1, 2)
    c = c / c.norm(dim=-1, keepdim=True)
    x = 0.5 * normal((H, nb, block, D), seed + 1) + gain * c[:, :, None, :]
    return x.reshape(1, H, nb * block, D)


def band_neighbours(n: int) -> torch.Tensor:
    """A small synthetic adjacency (self, +-1, -2, +5) standing in for the gilbert matrix."""
    nbr = torch.zeros(n, n, dtype=torch.bool)
    idx = torch.arange(n)
    for d in (-2, -1, 0, 1, 5):
        j = idx + d
        ok = (j >= 0) & (j < n)
        nbr[idx[ok], j[ok]] = True
    return nbr


def uniform_mask(shape, density: float, seed: int) -> torch.Tensor:
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
        u = (_splitmix64(idx) >> np.uint64(40)).astype(np.int64)
    return torch.from_numpy((u < int(density * (1 << 24))).reshape(shape))


ATTENTION_CASES = [
    # name, H, n_img, n_txt, seqlen, text_amp, seed
    ("hy", 2, 3, 2, 3 * 128 + 180, 0.431, 11),   # text keys amplified + padded text masked
    ("wan", 2, 4, 0, 4 * 128 - 8, 0.0, 12),      # ragged tail: 8 zero rows, no text blocks
    ("amp0", 1, 2, 1, 2 * 128 + 128, 0.0, 13),
]


def attention_case(name):
    for n, H, n_img, n_txt, seqlen, amp, seed in ATTENTION_CASES:
        if n != name:
            continue
        nb = n_img + n_txt
        S = nb * 128
        q = normal((1, H, S, 128), seed * 10 + 0).half()
        k = normal((1, H, S, 128), seed * 10 + 1).half()
        v = normal((1, H, S, 128), seed * 10 + 2).half()
        if n == "wan":  # zero padding rows like F.pad in wan/…:448-451
            q[:, :, seqlen:] = 0
            k[:, :, seqlen:] = 0
            v[:, :, seqlen:] = 0
        mask = uniform_mask((1, H, n_img, nb), 0.45, seed * 10 + 3)
        for i in range(n_img):
            mask[:, :, i, i] = True
        mask[..., n_img:] = True
        return dict(H=H, n_img=n_img, n_txt=n_txt, seqlen=seqlen, amp=amp, q=q, k=k, v=v, mask=mask)
    raise KeyError(name)


MASK_CASES = [
    # name, variant, H, n_img, n_txt, top_k, p, first_frame, gain, use_nbr, seed
    ("hy_u", "hyvideo", 2, 24, 2, 6, 0.3, 0, 0.0, True, 21),
    ("hy_p", "hyvideo", 2, 24, 2, 6, 0.3, 0, 3.0, True, 22),
    ("hy_nonbr", "hyvideo", 1, 16, 2, 4, 0.5, 0, 2.0, False, 23),
    ("wan_p", "wan", 2, 32, 0, 16, 0.9, 3, 3.0, True, 24),
    ("wan_u", "wan", 1, 20, 0, 10, 0.8, 2, 0.0, True, 25),
]


def mask_case(name):
    for n, variant, H, n_img, n_txt, top_k, p, ff, gain, use_nbr, seed in MASK_CASES:
        if n != name:
            continue
        nb = n_img + n_txt
        if gain > 0:
            q = peaky(H, nb, 128, gain, seed * 10).bfloat16()
            k = peaky(H, nb, 128, gain, seed * 10 + 5).bfloat16()
        else:
            q = normal((1, H, nb * 128, 128), seed * 10).bfloat16()
            k = normal((1, H, nb * 128, 128), seed * 10 + 5).bfloat16()
        nbr = band_neighbours(n_img) if use_nbr else None
        return dict(variant=variant, H=H, n_img=n_img, n_txt=n_txt, top_k=top_k, p=p, ff=ff, q=q,
                    k=k, nbr=nbr)
    raise KeyError(name)


PROLOGUE_CASE = dict(grid=(4, 8, 8), T=64, H=3, seed=41)  # L = 256 image tokens + 64 text tokens


def prologue_case():
    """img/txt fused-QKV activations [1, tokens, 3*H*128] bf16 and norm weights [128] bf16."""
    c = PROLOGUE_CASE
    t, h, w = c["grid"]
    L, T, H = t * h * w, c["T"], c["H"]
    img = (1.5 * normal((1, L, 3 * H * 128), c["seed"])).bfloat16()
    txt = (0.7 * normal((1, T, 3 * H * 128), c["seed"] + 1)).bfloat16()
    ws = [(1.0 + 0.2 * normal((128,), c["seed"] + 2 + i)).bfloat16() for i in range(4)]
    return dict(L=L, T=T, H=H, grid=c["grid"], img=img, txt=txt, w_img_q=ws[0], w_img_k=ws[1],
                w_txt_q=ws[2], w_txt_k=ws[3])


WAN_PROLOGUE_CASE = dict(grid=(3, 4, 6), pad=8, H=2, seed=61)  # 72 grid tokens + 8 padding tokens


def wan_prologue_case():
    """q/k projection outputs [1, L, H*128] bf16, fp32 norm weights [H*128]."""
    c = WAN_PROLOGUE_CASE
    f, h, w = c["grid"]
    L = f * h * w + c["pad"]
    C = c["H"] * 128
    xq = (1.3 * normal((1, L, C), c["seed"])).bfloat16()
    xk = (0.9 * normal((1, L, C), c["seed"] + 1)).bfloat16()
    wq = 1.0 + 0.2 * normal((C,), c["seed"] + 2)
    wk = 1.0 + 0.2 * normal((C,), c["seed"] + 3)
    n = f * h * w
    # a fixed pseudo-random permutation standing in for hilbert_order
    remap = torch.from_numpy(np.argsort(_splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(99)), kind="stable").astype(np.int64))
    return dict(L=L, H=c["H"], grid=c["grid"], xq=xq, xk=xk, wq=wq, wk=wk, remap=remap)


OPERATOR_CASES = [
    # name, variant, H, img_tokens, txt_tokens, txt_valid, top_k, p, text_amp, shape_xfuse, seed
    ("hy_op", "hyvideo", 2, 6 * 128, 256, 150, 2, 0.3, 0.431, False, 61),
    ("i2v_op", "hyvideo_i2v", 2, 5 * 128, 400, 300, 2, 0.3, 0.0, True, 62),
]


def operator_case(name):
    """fp16 [B,S,H,D] inputs of one whole-operator call (reference block_sparse_attention)."""
    for n, variant, H, L, T, t_valid, top_k, p, amp, xfuse, seed in OPERATOR_CASES:
        if n != name:
            continue
        S = L + T
        nbk = (S + 127) // 128
        q = peaky(H, nbk, 128, 2.0, seed * 10)[:, :, :S].transpose(1, 2).contiguous().half()
        k = peaky(H, nbk, 128, 2.0, seed * 10 + 3)[:, :, :S].transpose(1, 2).contiguous().half()
        v = normal((1, S, H, 128), seed * 10 + 6).half()
        cu = torch.tensor([0, L + t_valid, S], dtype=torch.int32)
        text_blocks = 2 if variant == "hyvideo" else 4
        n_img = (S + 127) // 128 - text_blocks
        return dict(variant=variant, H=H, L=L, T=T, S=S, q=q, k=k, v=v, cu=cu, top_k=top_k, p=p,
                    amp=amp, xfuse=xfuse, text_blocks=text_blocks, n_img=n_img,
                    nbr=band_neighbours(n_img))
    raise KeyError(name)
