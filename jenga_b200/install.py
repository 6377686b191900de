"""Makes the UNMODIFIED reference scripts (jenga_hyvideo.py, jenga_hyvideo_multigpu.py,
jenga_hyi2v.py, jenga_wan.py) import the B200 path instead of Triton / FlashAttention / the
Python gilbert module.

The reference binds names at import time
  from .attention_block_triton_diffres import block_sparse_attention
      (hyvideo/modules/models_mul_block_gc_ha_multigpu.py:25, xdit_ring_atten.py:16,
       hyvideo_i2v/modules/models_mul.py:22, wan/modules/model_mul.py:9)
  from flash_attn import flash_attn_func / flash_attn.flash_attn_interface...
  from gilbert import gilbert_mapping, ...
so the replacement must be in sys.modules BEFORE those imports run:

    python -m jenga_b200.install jenga_hyvideo.py --video-size 720 1280 ...      # launcher form
or  import jenga_b200.install; jenga_b200.install.install()                        # in-process
"""
from __future__ import annotations

import runpy
import sys
import types


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__jenga_b200__ = True
    return m


def install(flash_attn: bool = True, gilbert: bool = True, ulysses: bool = True, fused_blocks: bool = True,
            gather: bool = True) -> list[str]:
    """Pre-seeds sys.modules; returns the names it installed.
    fused_blocks: also patch the reference block classes right after they are imported so that
    RMSNorm + RoPE + cat + pooling run as the fused prologue kernel (jenga_b200/blocks.py).
    gather: route the token reorder `x[:, hilbert_order]` to the gather kernel."""
    from . import attention as A
    from . import flash_attn_shim as F
    from . import gilbert as G
    from . import ulysses as U

    def variant(name, text_blocks, p):
        def block_sparse_attention(query, key, value, top_k, block_size_M=128, block_size_N=128,
                                   cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None,
                                   max_seqlen_kv=None, text_blocks=text_blocks, text_amp=0.0,
                                   block_neighbor_list=None, shape_xfuse=False, p_remain_rates=p,
                                   first_frame_blocks=0):
            return A.block_sparse_attention_variant(
                name, query, key, value, top_k, block_size_M, block_size_N, cu_seqlens_q,
                cu_seqlens_kv, max_seqlen_q, max_seqlen_kv, text_blocks, text_amp,
                block_neighbor_list, shape_xfuse, p_remain_rates, first_frame_blocks)
        block_sparse_attention.__doc__ = f"jenga_b200 drop-in ({name} variant)"
        return block_sparse_attention

    done = []
    for pkg, name, tb, p in (("hyvideo", "hyvideo", 2, 0.5), ("hyvideo_i2v", "hyvideo_i2v", 4, 0.5),
                             ("wan", "wan", 0, 0.9)):
        fn = variant(name, tb, p)
        modname = f"{pkg}.modules.attention_block_triton_diffres"
        sys.modules[modname] = _module(modname, block_sparse_attention=fn,
                                       block_sparse_attention_combined=fn)
        done.append(modname)
    if ulysses:
        import os

        class xFuserLongContextAttention:
            """Same constructor surface as the reference class (all arguments ignored: ring degree
            is 1 in every Jenga script, xdit_ring_atten.py:24-59) and the same forward contract.
            The exchange runs fused over NVLink peer memory (UlyssesFusedAttention); if symmetric
            memory cannot be set up on this box — or JENGA_ULYSSES=nccl — all ranks agree on the
            NCCL all-to-all implementation.  Both are GPU paths of this library."""
            _impl = None

            def __init__(self, *a, **k):
                pass

            @classmethod
            def _get(cls):
                if cls._impl is None:
                    import torch
                    import torch.distributed as dist
                    want_fused = os.environ.get("JENGA_ULYSSES", "fused") != "nccl"
                    ok = 0
                    if want_fused and dist.is_initialized() and dist.get_backend() == "nccl":
                        try:
                            import torch.distributed._symmetric_memory as symm
                            symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
                            ok = 1
                        except Exception:  # noqa: BLE001
                            ok = 0
                        t = torch.tensor([ok], device="cuda")
                        dist.all_reduce(t, op=dist.ReduceOp.MIN)
                        ok = int(t.item())
                    cls._impl = U.UlyssesFusedAttention() if ok else U.UlyssesCarvedAttention(group=None)
                return cls._impl

            def forward(self, attn, query, key, value, **kw):
                return self._get()(attn, query, key, value, **kw)

            __call__ = forward
        modname = "hyvideo.modules.xdit_ring_atten"
        sys.modules[modname] = _module(modname, xFuserLongContextAttention=xFuserLongContextAttention)
        done.append(modname)
    if flash_attn:
        iface = _module("flash_attn.flash_attn_interface", flash_attn_func=F.flash_attn_func,
                        flash_attn_varlen_func=F.flash_attn_varlen_func,
                        _flash_attn_forward=F._flash_attn_forward)
        top = _module("flash_attn", flash_attn_func=F.flash_attn_func,
                      flash_attn_varlen_func=F.flash_attn_varlen_func, flash_attn_interface=iface,
                      __version__=F.__version__)
        top.__path__ = []  # behave like a package for `import flash_attn.flash_attn_interface`
        sys.modules["flash_attn"] = top
        sys.modules["flash_attn.flash_attn_interface"] = iface
        done += ["flash_attn", "flash_attn.flash_attn_interface"]
    if gilbert:
        sys.modules["gilbert"] = G
        done.append("gilbert")
    if fused_blocks:
        from . import blocks
        blocks.install_block_hook()
        done.append("block-forward hook")
    if gather:
        from . import blocks
        blocks.install_gather_hook()
        done.append("gather hook")
    return done


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m jenga_b200.install <reference_script.py> [script args...]")
    installed = install()
    print(f"[jenga_b200] installed: {', '.join(installed)}", file=sys.stderr)
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
