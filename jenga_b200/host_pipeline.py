"""Host-resident q/k/v -> AttenCarve -> host-resident output, pipelined by head group.

The operator is independent per head (selection, masks and attention are all per (batch, head),
SURVEY §8e), so a caller whose activations live in pinned host memory does not have to pay
H2D + compute + D2H back to back: heads are split into groups, and on three CUDA streams

    copy-in stream :  H2D(q,k,v of group g+1)      (strided 2-D copies out of [B,S,H,D])
    compute stream :  pool + select + carved attention of group g
    copy-out stream:  D2H(out of group g-1)

run concurrently (PCIe is full duplex).  Same arguments and results as
`block_sparse_attention` on the full tensors — the only difference is where the tensors live.
"""
from __future__ import annotations

import torch

from ._lib import check, lib
from .attention import block_sparse_attention_variant


def _copy2d(dst, dst_pitch, src, src_pitch, width, rows, direction, stream):
    check(lib.jenga_copy2d_async(dst, dst_pitch, src, src_pitch, width, rows, direction, stream.cuda_stream),
          "copy2d_async")


class HostPipelinedAttention:
    """Reusable pipeline (device staging buffers are allocated once)."""

    def __init__(self, B, S, H, D, dtype=torch.bfloat16, device="cuda", groups=4, variant="hyvideo"):
        if H % groups:
            raise ValueError("heads must divide by the number of groups")
        self.B, self.S, self.H, self.D, self.groups, self.variant = B, S, H, D, groups, variant
        self.hg = H // groups
        self.dev = torch.device(device)
        self.dtype = dtype
        mk = lambda: torch.empty((B, S, self.hg, D), dtype=dtype, device=self.dev)  # noqa: E731
        # double-buffered inputs and outputs
        self.q = [mk(), mk()]
        self.k = [mk(), mk()]
        self.v = [mk(), mk()]
        self.o = [mk(), mk()]
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.ev_in = [torch.cuda.Event() for _ in range(groups)]
        self.ev_done = [torch.cuda.Event() for _ in range(groups)]
        self.ev_free_in = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free_out = [torch.cuda.Event(), torch.cuda.Event()]
        self._used_in = [False, False]      # the event of that buffer has been recorded at least once
        self._used_out = [False, False]
        self._first = True

    def finish(self) -> None:
        """Current stream waits for every copy-out issued so far (after `__call__(..., wait=False)` steps)."""
        torch.cuda.current_stream(self.dev).wait_stream(self.s_out)

    def __call__(self, q_host, k_host, v_host, out_host, top_k, wait: bool = True, **kw):
        """q_host/k_host/v_host/out_host: pinned [B,S,H,D] host tensors (out_host is written).
        kw: the operator's keyword arguments (cu_seqlens_q/kv on the device, text_blocks, ...).
        wait=False lets consecutive calls overlap (copy-in of the next call under this call's
        attention, this call's copy-out under the next call); `finish()` before reading out_host."""
        B, S, H, D, hg = self.B, self.S, self.H, self.D, self.hg
        es = q_host.element_size()
        for t in (q_host, k_host, v_host, out_host):
            if t.shape != (B, S, H, D) or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("host tensors must be pinned, contiguous [B,S,H,D]")
        row_src, row_dst, width, rows = H * D * es, hg * D * es, hg * D * es, B * S
        cur = torch.cuda.current_stream(self.dev)
        if self._first:
            self.s_in.wait_stream(cur)    # staging buffers were allocated on `cur`
            self.s_out.wait_stream(cur)
            self._first = False

        def stage_in(g):
            b = g & 1
            if self._used_in[b]:
                self.s_in.wait_event(self.ev_free_in[b])  # the last compute that read these buffers is done
            off = g * hg * D * es
            for dst, src in ((self.q[b], q_host), (self.k[b], k_host), (self.v[b], v_host)):
                _copy2d(dst.data_ptr(), row_dst, src.data_ptr() + off, row_src, width, rows, 0, self.s_in)
            self.ev_in[g].record(self.s_in)

        stage_in(0)
        for g in range(self.groups):
            b = g & 1
            if g + 1 < self.groups:
                stage_in(g + 1)
            cur.wait_event(self.ev_in[g])
            if self._used_out[b]:
                cur.wait_event(self.ev_free_out[b])  # the last D2H out of this buffer has drained it
            block_sparse_attention_variant(self.variant, self.q[b], self.k[b], self.v[b], top_k,
                                           shape_xfuse=True, out=self.o[b], **kw)
            self.ev_done[g].record(cur)
            self.ev_free_in[b].record(cur)
            self._used_in[b] = True
            self.s_out.wait_event(self.ev_done[g])
            _copy2d(out_host.data_ptr() + g * hg * D * es, row_src, self.o[b].data_ptr(), row_dst, width, rows, 1,
                    self.s_out)
            self.ev_free_out[b].record(self.s_out)
            self._used_out[b] = True
        if wait:
            cur.wait_stream(self.s_out)
        return out_host
