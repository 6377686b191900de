"""Step-skip residual cache (HunyuanVideo), TeaCache gate (Wan2.1) and CUDA-graph replay of the
block loop's hot path — SURVEY §8 a-13 / f-2.

Mirrors the state machines the reference keeps on the model object:
  jenga_hyvideo.py:28,128-179            non_skip_steps, enable_skip, start_stage, cnt, previous_residual
  jenga_hyvideo_multigpu.py:225-238      the same + an all-reduce(MIN) of the (deterministic) decision
  jenga_wan.py:595-648, :1085-1098       TeaCache: ret_steps, cutoff_steps, coefficients, even/odd state
All tensor arithmetic runs in libjenga_b200.so (csrc/stepcache.cu); there is no torch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib
from .attention import _require_cuda, _stream_ptr

NON_SKIP_STEPS = (0, 1, 2, 3, 4, 7, 10, 13, 16, 19, 22, 25, 26, 29, 32, 35, 38, 41, 43, 45, 46, 47, 49)

_CODE = {torch.bfloat16: _lib.JENGA_BF16, torch.float16: _lib.JENGA_F16, torch.float32: _lib.JENGA_F32}


def _flat(x: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(x)
    if not x.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    if x.dtype not in _CODE:
        raise ValueError(f"{what}: dtype must be bf16, f16 or f32")
    return x


def residual_apply(x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """`x += residual` in place (jenga_hyvideo.py:130, jenga_wan.py:631,:641)."""
    _flat(x, "x"), _flat(residual, "residual")
    if x.shape != residual.shape or x.dtype != residual.dtype:
        raise ValueError("x and residual must have the same shape and dtype")
    with torch.cuda.device(x.device):
        check(lib.jenga_residual_apply(x.data_ptr(), residual.data_ptr(), x.numel(), _CODE[x.dtype],
                                       _stream_ptr(x.device)), "residual_apply")
    return x


def residual_store(x_new: torch.Tensor, x_old: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """`previous_residual = x_new - x_old` (jenga_hyvideo.py:179).  `out` lets the cache reuse its
    buffer instead of allocating 0.7 GB per computed step."""
    _flat(x_new, "x_new"), _flat(x_old, "x_old")
    if x_new.shape != x_old.shape or x_new.dtype != x_old.dtype:
        raise ValueError("x_new and x_old must have the same shape and dtype")
    if out is None:
        out = torch.empty_like(x_new)
    elif out.shape != x_new.shape or out.dtype != x_new.dtype or not out.is_contiguous():
        raise ValueError("out must match x_new")
    with torch.cuda.device(x_new.device):
        check(lib.jenga_residual_store(x_new.data_ptr(), x_old.data_ptr(), out.data_ptr(), x_new.numel(),
                                       _CODE[x_new.dtype], _stream_ptr(x_new.device)), "residual_store")
    return out


class StepSkipCache:
    """The skip logic of ra_forward (jenga_hyvideo.py:120-179), as an object a transformer
    forward drives:

        if cache.should_calc():            # decides, advances nothing
            ori = img                      # no clone needed: the blocks never write in place
            img = run_blocks(img)
            cache.store(img, ori)
        else:
            cache.apply(img)               # img += previous_residual
        cache.step()                       # cnt += 1, wraps at num_steps (:221-223)
    """

    def __init__(self, num_steps: int = 50, non_skip_steps=NON_SKIP_STEPS, enable_skip: bool = True):
        self.num_steps = int(num_steps)
        self.non_skip_steps = frozenset(non_skip_steps)
        self.enable_skip = enable_skip
        self.cnt = 0
        self.start_stage = False
        self.previous_residual: torch.Tensor | None = None

    def should_calc(self, sync_group=None) -> bool:
        if not self.enable_skip:
            return True
        calc = (self.cnt in self.non_skip_steps) or self.start_stage
        if calc:
            self.start_stage = False
        # jenga_hyvideo_multigpu.py:232-236 all-reduces this flag; every rank evaluates the same
        # deterministic list, so the collective (and its host sync) is dropped unless asked for
        if sync_group is not None:
            import torch.distributed as dist
            t = torch.tensor([1 if calc else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=sync_group)
            calc = bool(t.item())
        if not calc and self.previous_residual is None:
            raise RuntimeError("skipped step before any computed step (the reference would raise AttributeError)")
        return calc

    def apply(self, img: torch.Tensor) -> torch.Tensor:
        return residual_apply(img, self.previous_residual)

    def store(self, img_new: torch.Tensor, img_old: torch.Tensor) -> None:
        buf = self.previous_residual
        if buf is not None and (buf.shape != img_new.shape or buf.dtype != img_new.dtype):
            buf = None   # stage switch changed the token count
        self.previous_residual = residual_store(img_new, img_old, out=buf)

    def step(self) -> None:
        self.cnt += 1
        if self.cnt == self.num_steps:
            self.cnt = 0


class TeaCache:
    """TeaCache gate + residual cache of teacache_forward (jenga_wan.py:595-648).  One gate state
    per parity of `cnt` (even = conditional, odd = unconditional forward).  The decision is
    computed on the device; `gate()` returns it as a Python bool after waiting for ONE 4-byte flag
    in pinned host memory (the reference's `.cpu().item()` is the same synchronisation point)."""

    def __init__(self, coefficients, teacache_thresh: float, ret_steps: int, cutoff_steps: int, device="cuda"):
        if not 1 <= len(coefficients) <= 8:
            raise ValueError("1..8 polynomial coefficients")
        self.coefficients = [float(c) for c in coefficients]
        self.teacache_thresh = float(teacache_thresh)
        self.ret_steps, self.cutoff_steps = int(ret_steps), int(cutoff_steps)
        self.cnt = 0
        self.stage_start = False
        self.device = torch.device(device)
        self._state = torch.zeros(2, dtype=torch.float64, device=self.device)     # accum even / odd
        self._rel = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._flag = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._flag_dev_ptr = self._mapped_ptr(self._flag)
        self._prev: list[torch.Tensor | None] = [None, None]
        self.previous_residual: list[torch.Tensor | None] = [None, None]
        self._event = torch.cuda.Event()

    @staticmethod
    def _mapped_ptr(pinned: torch.Tensor) -> int:
        # pinned allocations of the CUDA runtime are mapped into the device address space (UVA)
        return pinned.data_ptr()

    def gate(self, modulated_inp: torch.Tensor) -> bool:
        """should_calc for the current forward; also performs `previous_e0 = modulated_inp.clone()`."""
        _require_cuda(modulated_inp)
        x = modulated_inp.contiguous()
        if x.dtype != torch.float32:
            raise ValueError("modulated input must be fp32 (the reference asserts e/e0 are fp32)")
        par = self.cnt % 2
        force = self.cnt < self.ret_steps or self.cnt >= self.cutoff_steps or self.stage_start or self._prev[par] is None
        if self._prev[par] is None or self._prev[par].shape != x.shape or self._prev[par].dtype != x.dtype:
            self._prev[par] = torch.empty_like(x)
            force = True
        a = _lib.JengaTeaCacheArgs()
        a.cur, a.prev = x.data_ptr(), self._prev[par].data_ptr()
        a.dtype, a.n = _CODE[x.dtype], x.numel()
        for i, c in enumerate(self.coefficients):
            a.coeff[i] = c
        a.n_coeff, a.thresh = len(self.coefficients), self.teacache_thresh
        a.force, a.update_prev = int(bool(force)), 1
        a.state = self._state.data_ptr() + 8 * par
        a.rel_out = self._rel.data_ptr() + 8 * par
        a.flag = self._flag_dev_ptr + 4 * par
        with torch.cuda.device(x.device):
            check(lib.jenga_teacache_gate(C.byref(a), _stream_ptr(x.device)), "teacache_gate")
        self._event.record(torch.cuda.current_stream(x.device))
        self._event.synchronize()
        return bool(self._flag[par].item())

    def accumulated(self, parity: int) -> float:
        return float(self._state[parity].item())

    def apply(self, x: torch.Tensor) -> torch.Tensor:
        return residual_apply(x, self.previous_residual[self.cnt % 2])

    def store(self, x_new: torch.Tensor, x_old: torch.Tensor) -> None:
        par = self.cnt % 2
        buf = self.previous_residual[par]
        if buf is not None and (buf.shape != x_new.shape or buf.dtype != x_new.dtype):
            buf = None
        self.previous_residual[par] = residual_store(x_new, x_old, out=buf)

    def step(self) -> None:
        self.cnt += 1          # jenga_wan.py:660


def prores_switch(latents: torch.Tensor, noise_pred: torch.Tensor, latents_noise: torch.Tensor, size,
                  d_sigma: float, sigma_next: float) -> torch.Tensor:
    """The latent update at a ProRes stage switch (pipeline_hunyuan_video_prores.py:721-731):
    `predict_x0_from_xt` -> `F.interpolate(..., size=size, mode="trilinear")` -> `add_noise_to_step`
    (scheduling_flow_match_discrete.py:258-299) in one kernel.  d_sigma = sigmas[-1] - sigmas[i],
    sigma_next = the sigma of timesteps[i+1].  Inputs are used as fp32 (the reference casts with
    .to(torch.float32)); returns fp32 [N, C, *size]."""
    _require_cuda(latents, noise_pred, latents_noise)
    lat, pred, noise = (t.to(torch.float32).contiguous() for t in (latents, noise_pred, latents_noise))
    if lat.dim() != 5 or lat.shape != pred.shape:
        raise ValueError("latents / noise_pred must be [N, C, T, H, W] of one shape")
    N, Cc, it, ih, iw = lat.shape
    ot, oh, ow = (int(v) for v in size)
    if tuple(noise.shape) != (N, Cc, ot, oh, ow):
        raise ValueError("latents_noise must be [N, C, *size]")
    out = torch.empty_like(noise)
    with torch.cuda.device(lat.device):
        check(lib.jenga_prores_switch(lat.data_ptr(), pred.data_ptr(), noise.data_ptr(), out.data_ptr(), N * Cc,
                                      it, ih, iw, ot, oh, ow, float(d_sigma), float(sigma_next),
                                      _stream_ptr(lat.device)), "prores_switch")
    return out


class GraphedHotPath:
    """CUDA-graph capture of a whole block loop's hot path (f-2): `fn(*static_inputs)` is run once
    eagerly (warm-up), captured once, then `replay()` launches every kernel of the loop with a
    single cudaGraphLaunch — removing the ~4 launches x 60 blocks of host work per forward.
    Inputs are static buffers: copy new activations INTO them (`copy_inputs`) before `replay()`;
    outputs are the tensors the captured call returned (overwritten by each replay)."""

    def __init__(self, fn, *static_inputs, warmup: int = 1):
        self.fn, self.inputs = fn, static_inputs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn(*static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the stream the warm-up ran on: per-stream state created there (the bulk prologue's
        # scratch for split blocks, csrc/prologue_bulk.cu) exists, so the capture takes the same kernels
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = fn(*static_inputs)

    def copy_inputs(self, *new_inputs) -> None:
        for dst, src in zip(self.inputs, new_inputs):
            if isinstance(dst, torch.Tensor) and src is not None and src is not dst:
                dst.copy_(src)

    def replay(self):
        self.graph.replay()
        return self.outputs
