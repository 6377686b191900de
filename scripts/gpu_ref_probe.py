#!/usr/bin/env python
"""GPU box: which tensor-core path does Triton's JIT give the reference kernel on sm_100?
Runs the unmodified reference operator once (tiny shape), dumps the cubin of
_triton_block_sparse_attn_fwd_kernel_onehot, disassembles it and counts MMA mnemonics.
Output: gpurun_out/ref_triton_isa.txt"""
import subprocess
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import refutil  # noqa: E402
from oracle import ref_loader  # noqa: E402

out = ROOT / "gpurun_out" / "ref_triton_isa.txt"
out.parent.mkdir(exist_ok=True)
wl = bench.workload("tiny", 0.7)
inp = bench.build_inputs(wl, torch.device("cuda", 0), heads=2)
op = ref_loader.operator("hyvideo")
refutil.reference_call(op, "hyvideo", inp["q"], inp["k"], inp["v"], top_k=inp["top_k"], cu=inp["cu"],
                       text_blocks=2, text_amp=0.0, nbr=inp["nbr"], p_remain=0.3)
torch.cuda.synchronize()
kern = op._triton_block_sparse_attn_fwd_kernel_onehot
lines = []
for dev, cache in kern.device_caches.items():
    for key, ck in cache[0].items():
        md = ck.metadata
        lines.append(f"kernel {ck.name if hasattr(ck, 'name') else ''} num_warps={getattr(md, 'num_warps', '?')} "
                     f"num_stages={getattr(md, 'num_stages', '?')} shared={getattr(md, 'shared', '?')} "
                     f"target={getattr(md, 'target', '?')}")
        ptx = ck.asm.get("ptx", "")
        for pat in ("tcgen05.mma", "tcgen05.ld", "mma.sync", "wgmma", "cp.async.bulk", "cp.async.ca", "cp.async.cg", "ld.global"):
            lines.append(f"  ptx {pat}: {ptx.count(pat)}")
        cubin = ck.asm.get("cubin")
        if cubin:
            with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
                f.write(cubin)
            r = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True)
            sass = r.stdout
            for pat in ("UTCHMMA", "UTCQMMA", "HMMA", "UTMALDG", "LDTM", "STTM", "LDGSTS", "LDG.E", "MUFU.EX2"):
                lines.append(f"  sass {pat}: {sass.count(pat)}")
out.write_text("\n".join(lines) + "\n")
print("\n".join(lines))
