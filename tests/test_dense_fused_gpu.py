"""f-1 (first slice): the fused elementwise chains against the reference's own ATen chains under
torch.autocast(bf16) — hyvideo/modules/models_mul_block_gc_ha_multigpu.py:196-199,295-315,409,499-500
with modulate_layers.py:31-68.  Tolerance: <= 1 bf16 ulp per element (4e-6 absolute where the
modulate sum cancels), >= 99.9 % bit-identical
(gate_residual: bit-exact; the others differ from ATen only in fp32 summation order / libm tanh)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _ulp_stats(got, ref):
    gi = got.view(torch.int16).to(torch.int32)
    ri = ref.view(torch.int16).to(torch.int32)
    # monotone integer image of the bf16 bit pattern
    gi = torch.where(gi < 0, -(gi & 0x7FFF), gi)
    ri = torch.where(ri < 0, -(ri & 0x7FFF), ri)
    d = (gi - ri).abs()
    return d.max().item(), (d == 0).float().mean().item()


def _modulate(x, shift, scale):                       # modulate_layers.py:31-49
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


@pytest.mark.parametrize("L,C", [(2048 + 37, 3072), (515, 1536), (1000, 5120)])
def test_ln_modulate_matches_autocast_chain(L, C):
    from jenga_b200 import dense as DN
    g = torch.Generator(device="cuda").manual_seed(L)
    x = (torch.randn(1, L, C, generator=g, device="cuda") * 1.7 + 0.3).bfloat16()
    shift = (0.3 * torch.randn(1, C, generator=g, device="cuda")).bfloat16()
    scale = (0.3 * torch.randn(1, C, generator=g, device="cuda")).bfloat16()
    norm = torch.nn.LayerNorm(C, elementwise_affine=False, eps=1e-6).cuda()
    lin = torch.nn.Linear(C, 8, bias=False).cuda().bfloat16()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        mid = _modulate(norm(x), shift, scale)
        assert mid.dtype == torch.float32                  # the chain the kernel restates
        ref = mid.to(torch.bfloat16)                       # the cast F.linear applies under autocast
        assert DN.ln_modulate_supported(x, norm, shift, scale)
        got = DN.ln_modulate(x, shift, scale, eps=norm.eps)
        assert torch.equal(lin(mid), lin(ref))             # i.e. the Linear sees exactly `ref`
    mx, same = _ulp_stats(got, ref)
    # n*t1 + shift cancels for some elements: there the fp32 sum is ~1e-7 * |operands| uncertain, which
    # is many ulps of a near-zero RESULT — bound those absolutely (operands are O(1))
    d = (got.float() - ref.float()).abs()
    bad = (d > 2.0 ** -7 * ref.float().abs() + 4e-6).sum().item()   # 1 bf16 ulp is up to 2^-7 relative
    print(f"\n[ln_modulate {L}x{C}] max ulp distance {mx}, identical {same:.5f}, beyond 1 ulp / 4e-6: {bad}")
    assert bad == 0 and same >= 0.999
    # strided input rows (a slice of a wider tensor)
    wide = torch.zeros(1, L, C + 64, dtype=torch.bfloat16, device="cuda")
    wide[:, :, :C] = x
    got2 = DN.ln_modulate(wide[:, :, :C], shift, scale, eps=1e-6)
    assert torch.equal(got2, got)


def test_gate_residual_is_bit_exact():
    from jenga_b200 import dense as DN
    g = torch.Generator(device="cuda").manual_seed(3)
    L, C = 3001, 3072
    x = torch.randn(1, L, C, generator=g, device="cuda").bfloat16()
    y = torch.randn(1, L, C, generator=g, device="cuda").bfloat16()
    gate = torch.randn(1, C, generator=g, device="cuda").bfloat16()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = x + y * gate.unsqueeze(1)                    # x + apply_gate(y, gate), modulate_layers.py:52-68
    assert DN.gate_residual_supported(x, y, gate)
    assert torch.equal(DN.gate_residual(x, y, gate), ref)
    # y as a strided slice
    wide = torch.randn(1, L, C + 128, generator=g, device="cuda").bfloat16()
    with torch.no_grad():
        ref2 = x + wide[:, :, 64:64 + C] * gate.unsqueeze(1)
    assert torch.equal(DN.gate_residual(x, wide[:, :, 64:64 + C], gate), ref2)


def test_gelu_tanh_into_strided_buffer():
    from jenga_b200 import dense as DN
    g = torch.Generator(device="cuda").manual_seed(4)
    L, Cq, Cm = 2000, 384, 1536
    lin1 = (2.5 * torch.randn(1, L, Cq + Cm, generator=g, device="cuda")).bfloat16()
    act = torch.nn.GELU(approximate="tanh")
    ref = act(lin1[:, :, Cq:])
    cat = torch.zeros(1, L, 128 + Cm, dtype=torch.bfloat16, device="cuda")
    DN.gelu_tanh_into(lin1[:, :, Cq:], cat[:, :, 128:])
    mx, same = _ulp_stats(cat[:, :, 128:].contiguous(), ref)
    print(f"\n[gelu_tanh] max ulp {mx}, identical {same:.6f}")
    assert mx <= 1 and same >= 0.999
    assert (cat[:, :, :128] == 0).all()
