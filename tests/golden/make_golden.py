#!/usr/bin/env python
"""Generates tests/golden/*.npz|json by executing the UNMODIFIED reference in the build container.

Run here (CPU container, /root/reference mounted):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed fixtures.

What is executed, and the two environment shims it needs (the reference files themselves are
imported by path and not edited):

* gilbert.py — imported with stub `matplotlib` modules (its plotting imports are absent here).
* hyvideo|wan/modules/attention_block_triton_diffres.py
    - `_triton_block_sparse_attn_fwd_kernel_onehot` runs under TRITON_INTERPRET=1 (numpy).  The
      interpreter passes Python-float kernel arguments through as Python scalars, which would
      make `q * qk_scale` a low-precision multiply; the COMPILED kernel receives them as fp32
      scalars (triton mangles float -> fp32), so `_implicit_cvt` is wrapped to box floats as
      fp32 tensors.  The interpreter's bfloat16 support is broken in triton 3.6 (garbage
      output), so attention fixtures are fp16 — same kernel source, `dtype=tl.float16`.
      `torch.cuda.device` is replaced by a null context (the launcher enters it on a CPU tensor).
    - `_build_block_index_with_importance_optimized` runs on CPU tensors as is.  The pipelines
      call it under torch.autocast("cuda", bfloat16) (pipeline_hunyuan_video_prores.py:663-665,
      jenga_wan.py:135) whose fp32 op list promotes softmax and cumsum; CPU autocast does not,
      so torch.softmax / torch.cumsum are wrapped to upcast to fp32 for the duration of the call.
    - `block_sparse_attention` (the whole operator, hyvideo and hyvideo_i2v variants): the glue,
      the mask builder and the Triton kernel run as above; the one third-party call on the path,
      `flash_attn.flash_attn_func` for the text rows (:377), needs a GPU and is replaced by an
      fp32 softmax attention on the same tensors (`_cpu_flash_attn_func`).  The wan variant casts
      to bfloat16 internally (wan/…:456), which the interpreter cannot run, so it has no
      operator-level fixture (its mask builder and prologue do).
"""
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys
import types
from pathlib import Path

os.environ["TRITON_INTERPRET"] = "1"

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))


def _import(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sha16(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ----------------------------------------------------------------------------- gilbert
def gilbert_goldens():
    for m in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    ref = _import("ref_gilbert", REF / "gilbert.py")
    cases = {}
    full = {}
    grids = [
        # (name, t, h, w, sliced, keep_full_arrays)
        ("toy_4x6x8", 4, 6, 8, 0, True),
        ("toy_5x7x9", 5, 7, 9, 0, True),
        ("toy_sliced_3x5x4", 3, 5, 4, 1, True),
        ("toy_sliced_8x11x13", 8, 11, 13, 1, True),
        ("odd_7x3x1", 7, 3, 1, 0, True),
        ("hy_r0.5_32x22x40", 32, 22, 40, 0, False),
        ("hy_r0.75_32x33x60", 32, 33, 60, 0, False),
        ("hy720p_32x45x80", 32, 45, 80, 0, False),
        ("wan1.3B_turbo_21x22x39", 21, 22, 39, 1, False),
        ("wan1.3B_21x30x52", 21, 30, 52, 1, False),
        ("wan14B_21x45x80", 21, 45, 80, 1, False),
    ]
    for name, t, h, w, sliced, keep in grids:
        with contextlib.redirect_stdout(io.StringIO()):
            if sliced:
                l2h, h2l = ref.sliced_gilbert_mapping(t, h, w)
                nbr = ref.sliced_gilbert_block_neighbor_mapping(t, h, w)
            else:
                l2h, h2l = ref.gilbert_mapping(t, h, w)
                nbr = ref.gilbert_block_neighbor_mapping(t, h, w)
        l2h = np.asarray(l2h, dtype=np.int64)
        h2l = np.asarray(h2l, dtype=np.int64)
        nbr = nbr.numpy().astype(np.bool_)
        cases[name] = dict(t=t, h=h, w=w, sliced=sliced, l2h_sha=sha16(l2h), h2l_sha=sha16(h2l),
                           nbr_sha=sha16(nbr), nbr_popcount=int(nbr.sum()),
                           nbr_row_max=int(nbr.sum(1).max()), l2h_head=l2h[:8].tolist())
        if keep:
            full[name + "/l2h"] = l2h
            full[name + "/h2l"] = h2l
            full[name + "/nbr"] = nbr
        print("gilbert", name, cases[name]["l2h_sha"], cases[name]["nbr_popcount"], flush=True)
    (OUT / "gilbert.json").write_text(json.dumps(cases, indent=1))
    np.savez_compressed(OUT / "gilbert_small.npz", **full)


# ----------------------------------------------------------------------------- attention (a-9)
def _patch_interpreter():
    import triton.language as tl
    import triton.runtime.interpreter as ti
    orig = ti._implicit_cvt

    def cvt(arg):
        if isinstance(arg, float):  # compiled kernels see float args as fp32 scalars
            return tl.tensor(ti.TensorHandle(np.array([arg], dtype=np.float32), tl.float32),
                             tl.float32)
        return orig(arg)

    ti._implicit_cvt = cvt
    torch.cuda.device = lambda d: contextlib.nullcontext()


def attention_goldens():
    _patch_interpreter()
    import synth
    hy = _import("ref_hy_attn", REF / "hyvideo/modules/attention_block_triton_diffres.py")
    out = {}
    for name, *_ in synth.ATTENTION_CASES:
        c = synth.attention_case(name)
        seqlens = torch.tensor([c["seqlen"]], dtype=torch.int32)
        o = hy._triton_block_sparse_attention_onehot(
            c["q"][:, :, :c["n_img"] * 128].contiguous(), c["k"], c["v"], seqlens, c["mask"],
            128 ** -0.5, text_amp=c["amp"], text_block_start=c["n_img"])
        out[name + "/o"] = o.numpy()
        out[name + "/mask_sha"] = np.frombuffer(sha16(c["mask"].numpy()).encode(), dtype=np.uint8)
        print("attention", name, float(o.float().abs().mean()), flush=True)
    np.savez_compressed(OUT / "attention_fp16.npz", **out)


# ----------------------------------------------------------------------------- mask builder (a-8)
@contextlib.contextmanager
def _cuda_autocast_dtype_flow():
    sm, cs = torch.softmax, torch.cumsum
    torch.softmax = lambda x, dim=-1, **kw: sm(x.float(), dim=dim, **kw)
    torch.cumsum = lambda x, dim=-1, **kw: cs(x.float(), dim=dim, **kw)
    try:
        yield
    finally:
        torch.softmax, torch.cumsum = sm, cs


def mask_goldens():
    import synth
    mods = {"hyvideo": _import("ref_hy_attn2", REF / "hyvideo/modules/attention_block_triton_diffres.py"),
            "wan": _import("ref_wan_attn", REF / "wan/modules/attention_block_triton_diffres.py")}
    out = {}
    for name, *_ in synth.MASK_CASES:
        c = synth.mask_case(name)
        nb = c["n_img"] + c["n_txt"]
        kw = dict(text_start_block=c["n_img"], num_blocks=nb, prob_threshold=c["p"],
                  text_blocks=c["n_txt"], block_neighbor_list=c["nbr"])
        if c["variant"] == "wan":
            kw["first_frame_blocks"] = c["ff"]
        with _cuda_autocast_dtype_flow():
            m = mods[c["variant"]]._build_block_index_with_importance_optimized(
                c["q"][:, :, :c["n_img"] * 128], c["k"], c["top_k"], 128, 128, **kw)
        out[name + "/mask"] = m.numpy()
        print("mask", name, "row counts", m[0, 0].sum(-1)[:8].tolist(), flush=True)
    np.savez_compressed(OUT / "mask_builder.npz", **out)


# ----------------------------------------------------------------------------- prologue (a-5, a-6)
def prologue_goldens():
    """RMSNorm + apply_rotary_emb exactly as MMDoubleStreamBlock.forward chains them
    (models_mul_block_gc_ha_multigpu.py:200-241), using the reference modules imported by path
    and its own RoPE table builder (theta 256, dims [16,56,56])."""
    import synth
    from einops import rearrange
    norm = _import("ref_norm_layers", REF / "hyvideo/modules/norm_layers.py")
    pos = _import("ref_posemb_layers", REF / "hyvideo/modules/posemb_layers.py")
    c = synth.prologue_case()
    H = c["H"]
    cos, sin = pos.get_nd_rotary_pos_embed([16, 56, 56], list(c["grid"]), theta=256, use_real=True,
                                           theta_rescale_factor=1)
    l2h = torch.randperm(c["L"], generator=torch.Generator().manual_seed(3))  # stands in for hilbert_order
    cos_g, sin_g = cos[l2h], sin[l2h]

    def mk(w):
        m = norm.RMSNorm(128, elementwise_affine=True, eps=1e-6, dtype=torch.bfloat16)
        with torch.no_grad():
            m.weight.copy_(w)
        return m

    iq, ik, iv = rearrange(c["img"], "B L (K H D) -> K B L H D", K=3, H=H)
    tq, tk, tv = rearrange(c["txt"], "B L (K H D) -> K B L H D", K=3, H=H)
    with torch.no_grad():
        iq = mk(c["w_img_q"])(iq).to(iv)
        ik = mk(c["w_img_k"])(ik).to(iv)
        iq, ik = pos.apply_rotary_emb(iq, ik, (cos_g, sin_g), head_first=False)
        tq = mk(c["w_txt_q"])(tq).to(tv)
        tk = mk(c["w_txt_k"])(tk).to(tv)
    q = torch.cat((iq, tq), dim=1)
    k = torch.cat((ik, tk), dim=1)
    v = torch.cat((iv, tv), dim=1)
    np.savez_compressed(OUT / "prologue.npz", q=q.view(torch.int16).numpy(), k=k.view(torch.int16).numpy(),
                        v_sha=np.frombuffer(sha16(v.contiguous().view(torch.int16).numpy()).encode(), dtype=np.uint8),
                        cos=cos.numpy(), sin=sin.numpy(), index=l2h.numpy())
    print("prologue", q.shape, float(q.float().abs().mean()), flush=True)


def wan_prologue_goldens():
    """WanRMSNorm + rope_apply (wan/modules/model_mul.py:74-90, :40-71, :145-151) from the
    reference file itself.  model_mul.py imports diffusers mixins and its sibling modules at
    import time; they are stubbed / resolved through a synthetic `wan.modules` package so that
    wan/__init__.py (which pulls the whole pipeline) is not executed."""
    import synth
    for name, attrs in (("diffusers", {}), ("diffusers.configuration_utils",
                                            {"ConfigMixin": object, "register_to_config": lambda f: f}),
                        ("diffusers.models", {}), ("diffusers.models.modeling_utils", {"ModelMixin": torch.nn.Module})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
    for name, path in (("wan", REF / "wan"), ("wan.modules", REF / "wan" / "modules")):
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m
    import importlib
    mm = importlib.import_module("wan.modules.model_mul")
    c = synth.wan_prologue_case()
    H, (f, h, w) = c["H"], c["grid"]
    d = 128
    freqs = torch.cat([mm.rope_params(1024, d - 4 * (d // 6)), mm.rope_params(1024, 2 * (d // 6)),
                       mm.rope_params(1024, 2 * (d // 6))], dim=1)        # wan/modules/model_mul.py (WanModel.__init__)
    grid_sizes = torch.tensor([[f, h, w]])
    out = {}
    for tag, x, wt in (("q", c["xq"], c["wq"]), ("k", c["xk"], c["wk"])):
        norm = mm.WanRMSNorm(H * d, eps=1e-6)
        with torch.no_grad():
            norm.weight.copy_(wt)
            y = norm(x).view(1, c["L"], H, d)
            r = mm.rope_apply(y, grid_sizes, freqs, c["remap"])
        assert r.dtype == torch.float32
        out[tag] = r.to(torch.bfloat16).view(torch.int16).numpy()   # the operator's cast (…diffres.py:456-463)
    out["freqs_sha"] = np.frombuffer(sha16(torch.view_as_real(freqs).numpy()).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / "wan_prologue.npz", **out)
    for name in ("wan", "wan.modules", "wan.modules.model_mul", "wan.modules.attention",
                 "wan.modules.attention_block_triton_diffres"):
        sys.modules.pop(name, None)
    print("wan_prologue", out["q"].shape, flush=True)


def gilbert_transpose_goldens():
    """transpose_gilbert_mapping (gilbert.py:274) and the transpose_order arguments of the four
    table builders, from the unmodified reference."""
    for m in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    ref = _import("ref_gilbert", REF / "gilbert.py")
    cases = {}
    for dims, order in (([3, 4, 6], [2, 1, 0]), ([4, 6, 8], [1, 0, 2]), ([5, 3, 7], [0, 2, 1]),
                        ([2, 5, 4], [1, 2, 0]), ([6, 9, 10], [2, 0, 1]), ([4, 4, 4], None)):
        t, h, w = dims
        with contextlib.redirect_stdout(io.StringIO()):
            l2h, h2l = ref.transpose_gilbert_mapping(dims, order)
            g = ref.gilbert_mapping(t, h, w, order)
            sl = ref.sliced_gilbert_mapping(t, h, w, order)
            nbr = ref.gilbert_block_neighbor_mapping(t, h, w, 16, order)
            snbr = ref.sliced_gilbert_block_neighbor_mapping(t, h, w, 16, order)
        key = "x".join(map(str, dims)) + "_" + ("none" if order is None else "".join(map(str, order)))
        cases[key] = dict(dims=dims, order=order, l2h=[int(i) for i in l2h], h2l=[int(i) for i in h2l],
                          gilbert_mapping_l2h_sha=sha16(np.asarray(g[0], dtype=np.int64)),
                          sliced_mapping_l2h_sha=sha16(np.asarray(sl[0], dtype=np.int64)),
                          nbr16_sha=sha16(nbr.numpy().astype(np.bool_)),
                          sliced_nbr16_sha=sha16(snbr.numpy().astype(np.bool_)))
        print("gilbert transpose", key, cases[key]["nbr16_sha"], flush=True)
    (OUT / "gilbert_transpose.json").write_text(json.dumps(cases))


# ----------------------------------------------------------------------------- whole operator (a-11)
def _cpu_flash_attn_func(q, k, v, causal=False, softmax_scale=None, **kw):
    """Stand-in for the third-party FlashAttention-2 call (flash_attn.flash_attn_func, text rows,
    …triton_diffres.py:377), which cannot run without a GPU: fp32 softmax attention on [B,S,H,D],
    result in the input dtype.  Everything else in the call below is the unmodified reference."""
    assert not causal
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * softmax_scale
    return (torch.softmax(s, dim=-1) @ vf).transpose(1, 2).to(q.dtype)


def operator_goldens():
    """reference block_sparse_attention (glue + mask builder + Triton kernel in the interpreter +
    the FA2 stand-in above) on fp16 inputs: HunyuanVideo variant with cu_seqlens, I2V variant with
    padding / trim / shape_xfuse."""
    _patch_interpreter()
    import synth
    mods = {"hyvideo": _import("ref_hy_op", REF / "hyvideo/modules/attention_block_triton_diffres.py"),
            "hyvideo_i2v": _import("ref_i2v_op", REF / "hyvideo_i2v/modules/attention_block_triton_diffres.py")}
    out = {}
    for name, *_ in synth.OPERATOR_CASES:
        c = synth.operator_case(name)
        m = mods[c["variant"]]
        m.flash_attn_func = _cpu_flash_attn_func
        with _cuda_autocast_dtype_flow():
            o = m.block_sparse_attention(
                c["q"], c["k"], c["v"], c["top_k"], cu_seqlens_q=c["cu"], cu_seqlens_kv=c["cu"],
                text_blocks=c["text_blocks"], text_amp=c["amp"], block_neighbor_list=c["nbr"],
                shape_xfuse=c["xfuse"], p_remain_rates=c["p"])
        out[name + "/o"] = o.numpy()
        print("operator", name, tuple(o.shape), float(o.float().abs().mean()), flush=True)
    np.savez_compressed(OUT / "operator_fp16.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gilbert", "gilbert_transpose", "attention", "mask", "prologue", "wan_prologue", "operator"]
    if "gilbert_transpose" in which:
        gilbert_transpose_goldens()
    if "operator" in which:
        operator_goldens()
    if "wan_prologue" in which:
        wan_prologue_goldens()
    if "prologue" in which:
        prologue_goldens()
    if "gilbert" in which:
        gilbert_goldens()
    if "attention" in which:
        attention_goldens()
    if "mask" in which:
        mask_goldens()
